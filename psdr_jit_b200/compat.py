"""A minimal stand-in for the slice of Dr.Jit that psdr_jit's README and tutorials use around the renderer:

    from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD, Array3f as Vector3fD ...
    P = FloatD(0.); drjit.enable_grad(P)
    sc.param_map["Mesh[0]"].set_transform(Matrix4fD([[1., 0., 0., P * 100.], ...]))
    img = integrator.renderD(sc, 0); drjit.set_grad(P, 1.0); drjit.forward_to(img); g = drjit.grad(img)

Dr.Jit itself is not a dependency of this repo (the kernels are hand-written CUDA).  The stand-in keeps host-side
values in numpy and tracks, for every value, its partial derivatives with respect to the leaves that
``enable_grad`` marked (forward mode only -- reverse mode goes through torch autograd, see psdr_jit_b200.renderD).
``install()`` registers the modules ``drjit``, ``drjit.cuda``, ``drjit.cuda.ad`` and ``drjit.scalar`` in
``sys.modules`` unless a real Dr.Jit is importable.  Reference: the Dr.Jit API used by README.md:45-107 and
tutorials/*.ipynb of psdr-jit."""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict

import numpy as np

_NEXT_LEAF = [1]


class Float:
    """drjit.cuda(.ad).Float: a float32 array (or scalar) with forward-mode partials d(value)/d(leaf)."""

    def __init__(self, value=0.0, _tan: Dict[int, np.ndarray] = None):
        if isinstance(value, Float):
            value, _tan = value.v, dict(value.tan)
        self.v = np.asarray(value, dtype=np.float32)
        self.tan: Dict[int, np.ndarray] = _tan or {}
        self.leaf = 0          # > 0: this variable is an AD leaf (enable_grad)
        self.grad_seed = 0.0   # set_grad

    # -- arithmetic with tangent propagation
    @staticmethod
    def _of(x):
        return x if isinstance(x, Float) else Float(x)

    def _tangents(self):
        return {self.leaf: np.float32(1.0)} if self.leaf else self.tan

    def _bin(self, other, f, da, db):
        o = Float._of(other)
        out = Float(f(self.v, o.v))
        t = {}
        for leaf, d in self._tangents().items():
            t[leaf] = t.get(leaf, 0) + da(self.v, o.v) * d
        for leaf, d in o._tangents().items():
            t[leaf] = t.get(leaf, 0) + db(self.v, o.v) * d
        out.tan = {k: np.asarray(v, np.float32) for k, v in t.items()}
        return out

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: 1.0, lambda a, b: 1.0)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, lambda a, b: 1.0, lambda a, b: -1.0)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: b, lambda a, b: a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b, lambda a, b: 1.0 / b, lambda a, b: -a / (b * b))
    def __radd__(self, o): return Float._of(o) + self
    def __rsub__(self, o): return Float._of(o) - self
    def __rmul__(self, o): return Float._of(o) * self
    def __rtruediv__(self, o): return Float._of(o) / self
    def __neg__(self): return self * -1.0
    def __float__(self): return float(self.v)
    def __len__(self): return int(self.v.size)
    def numpy(self): return np.array(self.v, dtype=np.float32)
    def __repr__(self): return "Float(%s)" % (self.v,)


Int = UInt64 = Float      # only used as array containers by the tutorials


class _Vec:
    """ArrayNf: N Float components."""
    n = 3

    def __init__(self, *comps):
        if len(comps) == 1 and not isinstance(comps[0], Float) and np.ndim(comps[0]) >= 1 and len(comps[0]) == self.n and np.ndim(comps[0]) == 1:
            comps = tuple(comps[0])
        if len(comps) == 1:
            comps = comps * self.n
        if len(comps) != self.n:
            raise TypeError("Array%df: expected %d components" % (self.n, self.n))
        self.c = [Float._of(x) for x in comps]

    def __getitem__(self, i): return self.c[i]
    def __setitem__(self, i, v): self.c[i] = Float._of(v)
    x = property(lambda s: s.c[0])
    y = property(lambda s: s.c[1])
    z = property(lambda s: s.c[2])

    def numpy(self):
        cols = np.broadcast_arrays(*[np.atleast_1d(c.v) for c in self.c])
        return np.stack(cols, axis=-1).astype(np.float32)

    def __psdr_value__(self):
        return self.numpy()

    def __psdr_tangents__(self):
        out = {}
        for i, c in enumerate(self.c):
            for leaf, d in c._tangents().items():
                m = out.setdefault(leaf, np.zeros(self.numpy().shape, np.float32))
                m[..., i] = d
        return out


class Array2f(_Vec):
    n = 2


class Array3f(_Vec):
    n = 3


class Matrix4f:
    """drjit Matrix4f built from nested lists whose entries are numbers or Floats."""

    def __init__(self, rows=None):
        self.m = [[Float(1.0 if i == j else 0.0) for j in range(4)] for i in range(4)]
        if rows is not None:
            if isinstance(rows, Matrix4f):
                rows = rows.m
            rows = list(rows)
            if len(rows) != 4 or any(len(list(r)) != 4 for r in rows):
                raise TypeError("Matrix4f: expected 4 x 4 entries")
            self.m = [[Float._of(x) for x in r] for r in rows]

    def __getitem__(self, ij):
        return self.m[ij[0]][ij[1]] if isinstance(ij, tuple) else self.m[ij]

    def numpy(self):
        return np.array([[float(np.ravel(x.v)[0]) for x in r] for r in self.m], dtype=np.float32)

    def __psdr_value__(self):
        return self.numpy()

    def __psdr_tangents__(self):
        out = {}
        for i in range(4):
            for j in range(4):
                for leaf, d in self.m[i][j]._tangents().items():
                    out.setdefault(leaf, np.zeros((4, 4), np.float32))[i, j] = float(np.ravel(d)[0])
        return out

    def __matmul__(self, o):
        a, b = self, Matrix4f(o)
        r = Matrix4f()
        for i in range(4):
            for j in range(4):
                acc = Float(0.0)
                for k in range(4):
                    acc = acc + a.m[i][k] * b.m[k][j]
                r.m[i][j] = acc
        return r


LEAVES: Dict[int, Float] = {}


def enable_grad(*xs):
    for x in xs:
        if not isinstance(x, Float):
            raise TypeError("enable_grad: only Float leaves are supported by the stand-in")
        x.leaf = _NEXT_LEAF[0]
        x.tan = {}
        LEAVES[x.leaf] = x
        _NEXT_LEAF[0] += 1


def set_grad(x, value):
    if not isinstance(x, Float) or not x.leaf:
        raise RuntimeError("set_grad: not an AD leaf (call enable_grad first)")
    x.grad_seed = float(value)


def detach(x):
    if isinstance(x, Float):
        return Float(x.v)
    return x


def forward_to(*imgs):
    """Forward-mode traversal up to the rendered images: one more kernel pass with the seeded tangents, replaying the
    sample streams of the renderD call that produced each image."""
    for img in imgs:
        if not hasattr(img, "_forward"):
            raise TypeError("forward_to: expected an image returned by renderD")
        img._forward({leaf: x.grad_seed for leaf, x in LEAVES.items() if x.grad_seed != 0.0})


def forward(x):
    raise NotImplementedError("drjit.forward(x): use set_grad(x, 1); forward_to(img)")


def grad(img):
    if isinstance(img, Float):
        return Float(img.grad_seed)
    g = getattr(img, "_grad", None)
    if g is None:
        raise RuntimeError("grad: no gradient (call forward_to(img) first)")
    return g


def eval(*a, **k):        # noqa: A001  (mirrors drjit.eval)
    return None


def sync_thread():
    try:
        import torch
        torch.cuda.synchronize()
    except Exception:
        pass


@contextlib.contextmanager
def suspend_grad(*a, **k):
    yield


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__psdr_stand_in__ = True
    return m


def install(force: bool = False) -> bool:
    """Registers the stand-in as `drjit` unless the real package imports.  Returns True if the stand-in is active."""
    if not force:
        try:
            import drjit  # noqa: F401
            return bool(getattr(sys.modules["drjit"], "__psdr_stand_in__", False))
        except Exception:
            pass
    common = dict(Float=Float, Int=Int, UInt64=UInt64, Array2f=Array2f, Array3f=Array3f, Matrix4f=Matrix4f, PCG32=None)
    ad = _module("drjit.cuda.ad", **common)
    cuda = _module("drjit.cuda", ad=ad, **common)
    scalar = _module("drjit.scalar", **common)
    top = _module("drjit", cuda=cuda, scalar=scalar, enable_grad=enable_grad, set_grad=set_grad, forward_to=forward_to, forward=forward,
                  grad=grad, detach=detach, eval=eval, sync_thread=sync_thread, suspend_grad=suspend_grad)
    sys.modules.update({"drjit": top, "drjit.cuda": cuda, "drjit.cuda.ad": ad, "drjit.scalar": scalar})
    return True
