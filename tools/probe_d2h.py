#!/usr/bin/env python3
"""Device->host copy of an image-sized result into (a) a torch pinned tensor, (b) the page-locked /dev/shm buffer of
psdr_jit_b200.dist.SharedHostBuffer through tensor.copy_, (c) the same buffer through cudaMemcpyAsync (ctypes on the CUDA
runtime torch loaded): microseconds per copy for 1/8, 1/2 and all of a [2, 512*512, 3] float32 frame, host-timed around
copy + stream synchronize (what the end-to-end step pays)."""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psdr_jit_b200.dist import SharedHostBuffer

n = 2 * 512 * 512 * 3
dev = torch.device("cuda", 0)
src = torch.rand(n, device=dev)
pinned = torch.empty(n, dtype=torch.float32, pin_memory=True)
shared = SharedHostBuffer(n, 0, 1, "probe")
print("registered", shared.registered, "torch sees pinned:", shared.data.is_pinned(), flush=True)
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
stream = torch.cuda.current_stream()


def timeit(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


for frac in (8, 2, 1):
    m = n // frac

    def a():
        pinned[:m].copy_(src[:m], non_blocking=True); stream.synchronize()

    def b():
        shared.data[:m].copy_(src[:m], non_blocking=True); stream.synchronize()

    def c():
        rt.cudaMemcpyAsync(ctypes.c_void_p(shared.data.data_ptr()), ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(4 * m), 2, ctypes.c_void_p(stream.cuda_stream))
        stream.synchronize()

    def g():
        shared.gather(src, 1)

    res = {"pinned copy_": timeit(a), "shared copy_": timeit(b)}
    if rt is not None:
        res["shared cudaMemcpyAsync"] = timeit(c)
    if frac == 1:
        res["SharedHostBuffer.gather (world 1)"] = timeit(g)
    print("%.2f MB:" % (4 * m / 1e6), {k: round(v, 1) for k, v in res.items()}, flush=True)
