"""CPU tests of the drop-in boundary: libeidola.so loads, exports every symbol include/eidola.h declares, fails
loudly (EID_ERR_CUDA) instead of falling back when there is no GPU, and the host-side logic above the kernels
(glTF import, table builders, camera) matches the oracle bit for bit on a host-only scene (EID_DEVICE_NONE)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import eidola_b200 as eid
from eidola_b200 import abi, scenes

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    return eid.lib().eid_device_count() == 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eidola.h")).read()
    declared = set(re.findall(r"EID_API\s+[\w\s\*]+?\b(eid_\w+)\s*\(", hdr))
    assert declared == set(eid.EXPORTS), declared ^ set(eid.EXPORTS)
    L = eid.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.eid_version() >= 100


def test_header_cites_the_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "eidola.h")).read()
    for cite in ("renderer.cpp:154-206", "scene.cpp:57-125", "scene.cpp:777-826", "accelstruct.cpp:55-162", "renderer.cpp:209-225"):
        assert cite in hdr


def test_abi_struct_layouts():
    assert C.sizeof(abi.SceneCamera) == 336 and C.sizeof(abi.RtxState) == 100
    assert abi.RtxState.size.offset == 48
    s = abi.default_rtx_state(1920, 1080)      # sample_example.hpp:154-184
    assert (s.maxDepth, s.modulate, s.ReSTIRState, s.RISSampleNum, s.reservoirClamp, s.MIS, s.denoise) == (4, 1, abi.eTemporal, 4, 80, 1, 1)
    assert abs(s.environmentProb - 0.25) < 1e-7 and abs(s.sigLuminDirect - 0.4) < 1e-7 and abs(s.sigDepthIndirect - 1.0) < 1e-7


def test_no_cpu_fallback_without_gpu():
    if not _no_gpu():
        pytest.skip("a GPU is visible here")
    h = C.c_void_p()
    rc = eid.lib().eid_scene_create(C.byref(h), 0)
    assert rc == -5 and b"no CPU fallback" in eid.lib().eid_last_error()      # EID_ERR_CUDA
    s = eid.Scene(device=-1)
    s.load_arrays(scenes.cube_scene())
    with pytest.raises(eid.EidolaError, match="no CPU fallback"):
        eid.AccelStructure().create(s)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(eid.pkg, "_lib", None)
    monkeypatch.setattr(eid.pkg, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(eid.EidolaError, match="no CPU fallback"):
        eid.pkg.lib()


@pytest.mark.parametrize("maker", [scenes.cube_scene, scenes.cornell_scene, scenes.small_room])
def test_host_tables_match_oracle(maker, tmp_path):
    """Scene::load's table builders (scene.cpp:179-448, 700-772): arrays path and glTF round trip, bit-exact."""
    arrays = maker()
    o = ol.OracleScene()
    o.load_arrays(arrays)
    p = eid.Scene(device=-1)
    p.load_arrays(arrays)
    g = eid.Scene(device=-1)
    g.load(scenes.write_gltf(arrays, str(tmp_path / "scene.gltf")))
    e = eid.Scene(device=-1)
    e.load(scenes.write_gltf(arrays, str(tmp_path / "embedded.gltf"), embed=True))
    io = o.info()
    for s in (p, g, e):
        ip = s.info()
        for f in ("primMeshCount", "nodeCount", "materialCount", "puncLightCount", "trigLightCount", "triangleInstances",
                  "trigLightWeight", "puncLightWeight"):
            assert getattr(io, f) == getattr(ip, f), f
        for t in (abi.TABLE_MATERIALS, abi.TABLE_PUNC_LIGHTS, abi.TABLE_TRIG_LIGHTS, abi.TABLE_LIGHT_INFO, abi.TABLE_INSTANCE_DATA):
            assert o.table(t).tobytes() == s.table(t).tobytes(), t
        for pm in range(io.primMeshCount):
            assert o.table(abi.TABLE_VERTICES, pm).tobytes() == s.table(abi.TABLE_VERTICES, pm).tobytes()
            assert o.table(abi.TABLE_INDICES, pm).tobytes() == s.table(abi.TABLE_INDICES, pm).tobytes()
    # glTF camera node -> eye / direction / fov survive the round trip
    cam = arrays.camera
    for s in (g,):
        s.update_camera(640, 360)
        o.set_lookat(cam["eye"], cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        o.update_camera(640, 360)
        a, b = o.table(abi.TABLE_CAMERA)[:48], s.table(abi.TABLE_CAMERA)[:48]     # viewInverse, projInverse, projView
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_camera_update_matches_oracle_and_rolls_history():
    """Scene::updateCamera (scene.cpp:777-826): last* = previous call's values, lastPosition = previous eye."""
    arrays = scenes.cornell_scene()
    o = ol.OracleScene()
    o.load_arrays(arrays)
    p = eid.Scene(device=-1)
    p.load_arrays(arrays)
    prev = None
    for k in range(4):
        eye = (0.1 * k, 1.0, -3.6)
        for s in (o, p):
            s.set_lookat(eye, (0, 1, 0), (0, 1, 0), 45.0)
            s.update_camera(1920, 1080)
        assert o.table(abi.TABLE_CAMERA).tobytes() == p.table(abi.TABLE_CAMERA).tobytes()
        cam = p.get_camera()
        if prev is not None:
            assert bytes(cam.lastProjView) == bytes(prev.projView)
            assert np.allclose((cam.lastPosition.x, cam.lastPosition.y, cam.lastPosition.z), (0.1 * (k - 1), 1.0, -3.6))
        prev = cam
    # perspectiveVK: y flipped, constant sub-pixel shift folded into the projection (scene.cpp:783-787)
    proj_inv = np.array(cam.projInverse.m[:], np.float32).reshape(4, 4).T
    proj = np.linalg.inv(proj_inv.astype(np.float64))
    assert proj[1, 1] < 0 and abs(proj[0, 2] - 0.5 / 1920) < 1e-6 and abs(proj[1, 2] - 0.5 / 1080) < 1e-6


def test_gltf_import_features(tmp_path):
    """Importer semantics the table builders rely on: node hierarchy (TRS + matrix), uint16 indices, missing
    NORMAL/TANGENT/TEXCOORD/COLOR defaults, default material, KHR_lights_punctual, data: URIs."""
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint16)
    blob = pos.tobytes() + idx.tobytes()
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"children": [1, 2], "translation": [1, 2, 3]},
                  {"mesh": 0, "scale": [2, 2, 2]},
                  {"extensions": {"KHR_lights_punctual": {"light": 0}}, "translation": [0, 5, 0]}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                      {"bufferView": 1, "componentType": 5123, "count": 6, "type": "SCALAR"}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 48}, {"buffer": 0, "byteOffset": 48, "byteLength": 12}],
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
        "extensions": {"KHR_lights_punctual": {"lights": [{"type": "point", "intensity": 3.0, "color": [1, 0.5, 0.25]}]}},
    }
    path = tmp_path / "t.gltf"
    path.write_text(json.dumps(doc))
    s = eid.Scene(device=-1)
    s.load(str(path))
    info = s.info()
    assert (info.primMeshCount, info.nodeCount, info.materialCount, info.puncLightCount, info.triangleInstances) == (1, 1, 1, 1, 2)
    assert np.allclose(info.bboxMin[:], (1, 2, 3)) and np.allclose(info.bboxMax[:], (3, 4, 3))     # T(1,2,3) * S(2)
    v = s.table(abi.TABLE_VERTICES, 0)
    assert np.array_equal(s.table(abi.TABLE_INDICES, 0), idx.astype(np.uint32))
    n = np.zeros(3, np.float32)
    ol.lib().orc_decompress_unit_vec(int(v["normal"][0]), n.ctypes.data)
    assert np.allclose(n, (0, 0, 1), atol=1e-4)                                                 # generated face normal
    assert (v["color"] == 0xFFFFFFFF).all() and np.allclose(v["texcoord"], 0, atol=1e-30)       # defaults
    m = s.table(abi.TABLE_MATERIALS)[0]
    assert tuple(m["pbrBaseColorFactor"]) == (1, 1, 1, 1) and m["pbrMetallicFactor"] == 1 and m["ior"] == 1.5 and m["pbrBaseColorTexture"] == -1
    L = s.table(abi.TABLE_PUNC_LIGHTS)[0]
    assert np.allclose(L["position"], (1, 7, 3)) and L["intensity"] == 3.0 and L["impSamp"]["pdf"] == 1.0


def test_error_behaviour_host(tmp_path):
    """Status codes instead of asserts/exceptions (the reference asserts on load failure, scene.cpp:164-169)."""
    L = eid.lib()
    s = eid.Scene(device=-1)
    assert L.eid_scene_load_gltf(s._h, b"/nonexistent/file.gltf") == -2          # EID_ERR_IO
    bad = tmp_path / "bad.gltf"
    bad.write_text("{ not json")
    assert L.eid_scene_load_gltf(s._h, str(bad).encode()) == -3                  # EID_ERR_PARSE
    assert b"JSON" in L.eid_last_error()
    assert L.eid_scene_get_info(s._h, C.byref(abi.SceneInfo())) == -6            # EID_ERR_STATE: nothing loaded
    assert L.eid_scene_load_desc(s._h, None) == -1                               # EID_ERR_INVALID
    arrays = scenes.cube_scene()
    arrays.indices[0] = 999                                                      # index out of range
    assert L.eid_scene_load_desc(s._h, C.byref(arrays.desc())) == -1
    tex = scenes.cube_scene()
    tex.materials[0]["baseColorTexture"] = 3
    assert L.eid_scene_load_desc(s._h, C.byref(tex.desc())) == -1                # texture index beyond the texture table
    blend = scenes.cube_scene()
    blend.materials[0]["alphaMode"] = 2
    blend.materials[0]["baseColorFactor"] = (1, 1, 1, 0.5)
    assert L.eid_scene_load_desc(s._h, C.byref(blend.desc())) == 0               # BLEND instances load (stochastic alpha is implemented)
    assert L.eid_scene_table_bytes(None, 0, 0) == -1 and L.eid_renderer_buffer_bytes(None, 0) == -1
    assert L.eid_renderer_run(None, None, 0) == -1


def test_empty_and_degenerate_scenes():
    """Empty light tables keep one dummy record ("cannot be null", scene.cpp:349-351, 401-403)."""
    a = scenes.cube_scene()
    a.lights = []
    s = eid.Scene(device=-1)
    s.load_arrays(a)
    o = ol.OracleScene()
    o.load_arrays(a)
    assert s.info().puncLightCount == 0 and s.table(abi.TABLE_PUNC_LIGHTS).size == 1 and s.table(abi.TABLE_TRIG_LIGHTS).size == 1
    li = s.table(abi.TABLE_LIGHT_INFO)[0]
    assert li["trigSampProb"] == 0 and o.table(abi.TABLE_LIGHT_INFO).tobytes() == s.table(abi.TABLE_LIGHT_INFO).tobytes()


def test_synthetic_scene_generators_meet_the_contract():
    """triangle / light counts of the configs, material ids clear of the sky hash."""
    c2 = scenes.cornell_scene()
    assert c2.indices.size // 3 == 32
    s = eid.Scene(device=-1)
    s.load_arrays(c2)
    assert s.info().trigLightCount == 2
    for maker in (scenes.cube_scene, scenes.cornell_scene, scenes.small_room):
        a = maker()
        assert len(a.materials) < 200
        assert all(((i ^ (i >> 8)) & 0xff) != 0xff for i in range(len(a.materials)))
    c1 = scenes.cube_scene()
    assert c1.indices.size // 3 == 12 and c1.positions.shape[0] == 24


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    """host/eidola.hpp (the C++ mirror of Scene / AccelStructure / Renderer) + the headless harness build against the .so."""
    import shutil
    import subprocess
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    pkg = os.path.join(ROOT, "cis-565-final-vr-raytracer_b200")
    exe = str(tmp_path / "render_gltf")
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(pkg, "host", "render_gltf.cpp"), "-L", pkg, "-leidola", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=pkg)
    p = subprocess.run([exe, "/nonexistent.gltf", str(tmp_path / "o.pfm"), "64", "64", "1"], env=env, capture_output=True, text=True)
    if _no_gpu():
        assert p.returncode == 1 and "no CPU fallback" in p.stderr
    else:
        assert p.returncode == 1 and "load failed" in p.stderr


def test_hdr_environment_tables_and_loader(tmp_path):
    """HdrSampling host side (hdr_sampling.cpp:107-242): alias map, integral, average == oracle bit for bit; .hdr (RGBE) reader."""
    img = scenes.synthetic_sky()
    o = ol.OracleEnv(img)
    p = eid.HdrSampling(device=-1)
    p.set_pixels(img)
    assert o.accel().tobytes() == p.accel().tobytes()
    assert (o.get_integral(), o.get_average()) == (p.get_integral(), p.get_average())
    a = p.accel()
    n = a.size
    assert ((a["alias"] >= 0) & (a["alias"] < n)).all() and (a["q"] >= 0).all() and (a["q"] <= 1.0001).all()
    # the alias table reproduces the target distribution: P(i) = (q_i + sum_{j: alias_j = i} (1 - q_j)) / n == pdf_i * solid angle ~ importance
    mass = a["q"].astype(np.float64).copy()
    np.add.at(mass, a["alias"], 1.0 - a["q"].astype(np.float64))
    assert abs(mass.sum() / n - 1.0) < 1e-4
    # constant map: integral = 4*pi*value
    c = eid.HdrSampling(device=-1)
    c.set_pixels(np.full((8, 16, 4), 0.25, np.float32))
    assert abs(c.get_integral() - np.pi) < 1e-4
    for rle in (True, False):
        path = str(tmp_path / ("sky_%d.hdr" % rle))
        dec = scenes.write_radiance_hdr(path, img, rle=rle)
        q = eid.HdrSampling(device=-1)
        q.load_environment(path)
        assert np.array_equal(q.pixels(), dec) and q.size() == (img.shape[1], img.shape[0])
        assert np.abs(dec[..., :3] - img[..., :3]).max() <= 0.01 * img[..., :3].max()       # RGBE quantisation only
    L = eid.lib()
    h = C.c_void_p()
    assert L.eid_env_load_hdr(C.byref(h), -1, b"/nonexistent.hdr") == -2
    bad = tmp_path / "bad.hdr"
    bad.write_bytes(b"P6 not an hdr")
    assert L.eid_env_load_hdr(C.byref(h), -1, str(bad).encode()) == -3
    assert L.eid_env_create(C.byref(h), -1, None, 4, 4) == -1


def test_texture_table_semantics(tmp_path):
    """Scene::createTextureImages (scene.cpp:554-646): textured scene loads on the host path, tables match the oracle, a glTF
    that references undecoded images is refused until the host provides them (no PNG/JPEG decoder in this build)."""
    arrays = scenes.textured_scene()
    o = ol.OracleScene()
    o.load_arrays(arrays)
    p = eid.Scene(device=-1)
    p.load_arrays(arrays)
    assert o.table(abi.TABLE_MATERIALS).tobytes() == p.table(abi.TABLE_MATERIALS).tobytes()
    m = p.table(abi.TABLE_MATERIALS)
    assert m["pbrBaseColorTexture"][0] == 0 and m["normalTexture"][0] == 2 and m["emissiveTexture"][4] == 3
    # glTF with an external image: refused, then accepted once the decoded texels are provided
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
           "images": [{"uri": "albedo.png"}], "samplers": [{"magFilter": 9728, "wrapS": 33071}],
           "textures": [{"source": 0, "sampler": 0}],
           "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}]}
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes()
    doc["meshes"] = [{"primitives": [{"attributes": {"POSITION": 0}, "material": 0}]}]
    doc["accessors"] = [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}]
    doc["bufferViews"] = [{"buffer": 0, "byteOffset": 0, "byteLength": 36}]
    doc["buffers"] = [{"byteLength": 36, "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]
    path = tmp_path / "tex.gltf"
    path.write_text(json.dumps(doc))
    s = eid.Scene(device=-1)
    L = eid.lib()
    assert L.eid_scene_load_gltf(s._h, str(path).encode()) == -4 and b"eid_scene_provide_image" in L.eid_last_error()
    s.provide_image(0, np.full((4, 4, 4), 200, np.uint8))
    s.load(str(path))
    assert s.table(abi.TABLE_MATERIALS)["pbrBaseColorTexture"][0] == 0
