"""Scene I/O (reference src/scene/scene_loader.cpp, src/core/bitmap_loader.cpp, src/shape/mesh.cpp:469):
host-side parsing only -- no GPU needed until configure()."""
import os

import numpy as np
import pytest

from tests.common import scenes

XML = '''<scene version="0.6.0">
 <sensor type="perspective"><float name="fov" value="60"/><float name="nearClip" value="1e-6"/><float name="far_clip" value="1e7"/>
  <transform name="to_world"><scale x="1" y="1" z="1"/><rotate y="1" angle="0"/><translate x="278" y="273" z="-800"/></transform>
  <sampler type="independent"><integer name="sample_count" value="4"/></sampler>
  <film type="hdrfilm"><integer name="width" value="48"/><integer name="height" value="32"/></film></sensor>
 <sensor type="perspective"><float name="fov" value="45"/>
  <transform name="toWorld"><lookat origin="278, 273, -800" target="278, 273, 0" up="0, 1, 0"/></transform></sensor>
 <bsdf type="diffuse" id="white"><rgb name="reflectance" value="0.95"/></bsdf>
 <bsdf type="diffuse" id="light"><rgb name="reflectance" value="0,0,0"/></bsdf>
 <bsdf type="diffuse" id="tex"><texture type="bitmap" name="reflectance"><string name="filename" value="tex.exr"/></texture></bsdf>
 <bsdf type="microfacet" id="mf"><rgb name="specularReflectance" value="0.2,0.9,0.9"/><rgb name="diffuse_reflectance" value="0.01"/>
   <float name="roughness" value="0.3"/></bsdf>
 <shape type="obj" id="lum"><string name="filename" value="cbox_luminaire.obj"/><ref id="light"/>
   <transform name="to_world"><translate y="-0.5"/></transform><emitter type="area"><rgb name="radiance" value="20,20,8"/></emitter></shape>
 <shape type="obj"><string name="filename" value="cbox_floor.obj"/><ref id="mf"/><boolean name="face_normals" value="true"/></shape>
 <shape type="obj"><string name="filename" value="cbox_back.obj"/><ref id="tex"/></shape>
</scene>'''


def _write_exr(path, w=6, h=4):
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    cv2 = pytest.importorskip("cv2")
    rgb = np.random.default_rng(3).random((h, w, 3), dtype=np.float32)
    if not cv2.imwrite(path, np.ascontiguousarray(rgb[:, :, ::-1])):
        pytest.skip("this OpenCV build cannot write OpenEXR")
    return rgb


def test_xml_scene_loader(tmp_path):
    import psdr_jit_b200 as psdr
    scenes.write_cbox_objs(str(tmp_path))
    rgb = _write_exr(str(tmp_path / "tex.exr"))
    (tmp_path / "scene.xml").write_text(XML)
    sc = psdr.Scene()
    sc.load_file(str(tmp_path / "scene.xml"), auto_configure=False)
    o = sc.opts
    assert (o.width, o.height, o.spp, o.sppe, o.sppse) == (48, 32, 4, 0, 0)               # scene_loader.cpp:245-252
    assert sc.num_sensors == 2 and sc.num_meshes == 3 and sc.get_num_emitters() == 1
    assert np.allclose(sc.param_map["Sensor[0]"].to_world, scenes.translate(278, 273, -800))
    cam1 = sc.param_map["Sensor[1]"]
    assert (cam1.near, cam1.far) == (pytest.approx(0.1), pytest.approx(1e4))                # defaults, scene_loader.cpp:267-271
    assert np.allclose(cam1.to_world[:3, :3], np.eye(3), atol=1e-6) and np.allclose(cam1.to_world[:3, 3], [278, 273, -800])
    assert np.allclose(sc.param_map["BSDF[id=white]"].reflectance, 0.95)                   # a short rgb repeats its last entry
    mf = sc.param_map["BSDF[id=mf]"]
    assert np.allclose(mf.specularReflectance, [0.2, 0.9, 0.9]) and np.allclose(mf.diffuseReflectance, 0.01) and float(mf.roughness) == pytest.approx(0.3)
    tex = sc.param_map["BSDF[id=tex]"].reflectance
    assert tex.resolution == (6, 4) and np.allclose(tex.data.reshape(4, 6, 3), rgb, atol=2e-3)   # half-float EXR
    lum = sc.param_map["Mesh[id=lum]"]
    assert lum.num_faces == 2 and lum.to_world[1, 3] == pytest.approx(-0.5)
    assert np.allclose(sc.param_map["Emitter[0]"].radiance, [20, 20, 8])
    assert sc.param_map["Mesh[1]"].use_face_normal and not sc.param_map["Mesh[2]"].use_face_normal
    assert sc.param_map["Mesh[2]"].vertex_uv is not None
    for bad, msg in ((XML.replace('type="perspective"', 'type="orthographic"', 1), "Unsupported sensor"),
                     (XML.replace('<ref id="mf"/>', '<ref id="nope"/>'), "Unknown BSDF id"),
                     (XML.replace('type="microfacet"', 'type="plastic"'), "Unsupported BSDF"), ("<scene", "XML parsing failed")):
        with pytest.raises(RuntimeError, match=msg):
            os.chdir(tmp_path)
            psdr.Scene().load_string(bad, auto_configure=False)


def test_exr_bitmaps_and_envmap(tmp_path):
    import psdr_jit_b200 as psdr
    rgb = _write_exr(str(tmp_path / "env.exr"), 8, 4)
    b3, b1 = psdr.Bitmap3fD(str(tmp_path / "env.exr")), psdr.Bitmap1fD(str(tmp_path / "env.exr"))
    assert b3.resolution == (8, 4) and np.allclose(b3.data, rgb.reshape(-1, 3), atol=2e-3)
    assert b1.data.shape == (32, 1) and np.allclose(b1.data[:, 0], rgb.reshape(-1, 3)[:, 0], atol=2e-3)   # channel 0 (bitmap.cpp:37-39)
    sc = psdr.Scene()
    sc.add_EnvironmentMap(str(tmp_path / "env.exr"), scenes.translate(0, 0, 0), 2.5)
    env = sc.param_map["Emitter[0]"]
    assert env.radiance.resolution == (8, 4) and float(env.scale) == 2.5
    with pytest.raises(RuntimeError, match="only allowed to have one envmap"):
        sc.add_EnvironmentMap(psdr.EnvironmentMap(str(tmp_path / "env.exr")))
    with pytest.raises(RuntimeError, match="Failed to load EXR"):
        psdr.Bitmap3fD(str(tmp_path / "missing.exr"))


def test_mesh_dump_roundtrip(tmp_path):
    import psdr_jit_b200 as psdr
    for m in (scenes.cbox_meshes()[2], scenes.icosphere(1, 2.0, (1.0, 2.0, 3.0))):         # with and without uv
        mesh = psdr.Mesh()
        mesh.load_raw(m.v, m.f, m.uv, m.fuv)
        mesh.to_world = scenes.translate(1.0, 2.0, 3.0)
        for raw in (False, True):
            p = str(tmp_path / ("%s_%d.obj" % (m.name, raw)))
            mesh.dump(p, raw)
            back = psdr.Mesh()
            back.load(p)
            want = m.v + (np.array([1.0, 2.0, 3.0], np.float32) if raw else 0.0)
            assert np.allclose(back.vertex_positions, want, rtol=2e-6, atol=1e-5) and np.array_equal(back.face_indices, m.f)
            assert (back.vertex_uv is None) == (m.uv is None)
            assert "vn " in open(p).read()
