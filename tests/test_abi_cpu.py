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
    assert L.eid_scene_load_gltf(s._h, str(path).encode()) == -4 and b"eid_scene_provide_image" in L.eid_last_error()   # albedo.png does not exist
    s.provide_image(0, np.full((4, 4, 4), 200, np.uint8))
    s.load(str(path))
    assert s.table(abi.TABLE_MATERIALS)["pbrBaseColorTexture"][0] == 0


def _png(img, color_type, depth=8, interlace=0, palette=None, trns=None, filters=None):
    """Minimal PNG writer for the decoder tests: img = (h, w, channels) array of samples (uint8 or uint16), any filter per row."""
    import struct, zlib
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]

    def pack_rows(sub):
        rows = []
        for r in sub:
            v = r.reshape(-1)
            if depth == 16:
                rows.append(v.astype(">u2").tobytes())
            elif depth == 8:
                rows.append(v.astype(np.uint8).tobytes())
            else:
                bits = "".join(format(int(x), "0%db" % depth) for x in v)
                bits += "0" * (-len(bits) % 8)
                rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
        return rows

    def filt(rows, bpp):
        out, prev = b"", bytes(len(rows[0])) if rows else b""
        for y, cur in enumerate(rows):
            ft = (filters[y % len(filters)] if filters else 0)
            line = bytearray(len(cur))
            for i in range(len(cur)):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pae = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                pred = [0, a, b, (a + b) >> 1, pae][ft]
                line[i] = (cur[i] - pred) & 255
            out += bytes([ft]) + bytes(line)
            prev = cur
        return out

    bpp = max(1, ch * depth // 8)
    if interlace:
        raw = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = img[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += filt(pack_rows(sub), bpp)
    else:
        raw = filt(pack_rows(img), bpp)

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    z = zlib.compress(raw, 6)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, interlace))
    if palette is not None:
        out += chunk(b"PLTE", bytes(palette))
    if trns is not None:
        out += chunk(b"tRNS", bytes(trns))
    out += chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    return out


def _gltf_with_image(tmp_path, name, image_entry, extra_buffers=()):
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes()
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
           "images": [image_entry], "textures": [{"source": 0}],
           "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "material": 0}]}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}] + [dict(buffer=1 + k, byteOffset=0, byteLength=len(b)) for k, b in enumerate(extra_buffers)],
           "buffers": [{"byteLength": 36, "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}] +
                      [{"byteLength": len(b), "uri": "data:application/octet-stream;base64," + base64.b64encode(b).decode()} for b in extra_buffers]}
    path = tmp_path / name
    path.write_text(json.dumps(doc))
    return str(path)


def test_png_images_are_decoded_by_the_importer(tmp_path):
    """tinygltf loads texture images through stb_image with 4 requested components (scene.cpp:144-159): the importer decodes PNG with the
    same conventions — every colour type / bit depth / filter, Adam7, palette + tRNS, 16-bit -> high byte, colour keys — from an external
    file, a data: URI and a bufferView."""
    import base64
    rng = np.random.default_rng(7)
    cases = []
    rgba = rng.integers(0, 256, (13, 11, 4), dtype=np.uint8)
    cases.append(("rgba8 all filters", _png(rgba, 6, filters=[0, 1, 2, 3, 4]), rgba))
    rgb = rng.integers(0, 256, (9, 17, 3), dtype=np.uint8)
    want = np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)])
    cases.append(("rgb8 paeth", _png(rgb, 2, filters=[4]), want))
    cases.append(("rgb8 interlaced", _png(rgb, 2, interlace=1, filters=[1, 3]), want))
    key = rgb[2, 3]
    wk = want.copy(); wk[(rgb == key).all(axis=2), 3] = 0
    cases.append(("rgb8 colour key", _png(rgb, 2, trns=[0, key[0], 0, key[1], 0, key[2]]), wk))
    g16 = rng.integers(0, 65536, (5, 7, 2), dtype=np.uint16)
    w16 = np.dstack([g16[..., 0] >> 8] * 3 + [g16[..., 1] >> 8]).astype(np.uint8)
    cases.append(("grey+alpha 16", _png(g16, 4, depth=16, filters=[2, 4]), w16))
    for depth, scale in ((1, 255), (2, 85), (4, 17)):
        g = rng.integers(0, 1 << depth, (6, 13), dtype=np.uint8)
        wg = np.dstack([(g * scale).astype(np.uint8)] * 3 + [np.full(g.shape, 255, np.uint8)])
        cases.append(("grey %d-bit" % depth, _png(g, 0, depth=depth, filters=[0, 2]), wg))
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    idx = rng.integers(0, 16, (8, 9), dtype=np.uint8)
    tr = [0, 128, 255, 7]
    wp = np.dstack([pal[idx], np.array(tr + [255] * 12, np.uint8)[idx]])
    cases.append(("palette 4-bit + tRNS", _png(idx, 3, depth=4, palette=pal.reshape(-1), trns=tr, interlace=1), wp))
    for k, (name, png, want) in enumerate(cases):
        if k % 3 == 0:
            (tmp_path / ("img%d.png" % k)).write_bytes(png)
            entry = {"uri": "img%d.png" % k}
            path = _gltf_with_image(tmp_path, "c%d.gltf" % k, entry)
        elif k % 3 == 1:
            path = _gltf_with_image(tmp_path, "c%d.gltf" % k, {"uri": "data:image/png;base64," + base64.b64encode(png).decode()})
        else:
            path = _gltf_with_image(tmp_path, "c%d.gltf" % k, {"bufferView": 1, "mimeType": "image/png"}, extra_buffers=[png])
        s = eid.Scene(device=-1)
        s.load(path)
        got = s.table(abi.TABLE_TEXELS, 0).reshape(want.shape)
        assert np.array_equal(got, want), name
    # a truncated / corrupt PNG is a parse error, not a crash
    bad = cases[0][1][:60]
    (tmp_path / "bad.png").write_bytes(bad)
    L = eid.lib()
    s = eid.Scene(device=-1)
    assert L.eid_scene_load_gltf(s._h, _gltf_with_image(tmp_path, "bad.gltf", {"uri": "bad.png"}).encode()) == -3 and b"PNG" in L.eid_last_error()
    # a JPEG (not decoded by the library) is refused until the host provides the texels
    (tmp_path / "x.jpg").write_bytes(b"\xff\xd8\xff\xe0" + bytes(32))
    assert L.eid_scene_load_gltf(s._h, _gltf_with_image(tmp_path, "jpg.gltf", {"uri": "x.jpg"}).encode()) == -4


def test_malformed_gltf_is_a_parse_error_not_a_crash(tmp_path):
    """Every index and byte range the importer takes from the file is validated (EID_ERR_PARSE = -3), JSON nesting is capped."""
    import base64, copy
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    idx = np.array([0, 1, 2], np.uint16)
    blob = pos.tobytes() + idx.tobytes() + bytes(2)
    good = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                          {"bufferView": 1, "componentType": 5123, "count": 3, "type": "SCALAR"}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 6}],
            "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    L = eid.lib()
    s = eid.Scene(device=-1)

    def rc(doc, name="m.gltf"):
        p = tmp_path / name
        p.write_text(doc if isinstance(doc, str) else json.dumps(doc))
        return L.eid_scene_load_gltf(s._h, str(p).encode())
    assert rc(good) == 0
    for mutate in (lambda d: d["meshes"][0]["primitives"][0].__setitem__("indices", 7),                 # accessor index out of range
                   lambda d: d["meshes"][0]["primitives"][0]["attributes"].__setitem__("POSITION", 9),
                   lambda d: d["accessors"][1].__setitem__("bufferView", 5),                           # bufferView out of range
                   lambda d: d["bufferViews"][1].__setitem__("buffer", 3),                             # buffer out of range
                   lambda d: d["accessors"][0].__setitem__("count", -3),                               # negative count
                   lambda d: d["accessors"][0].__setitem__("count", 2.5),                              # non-integral count
                   lambda d: d["accessors"][0].__setitem__("count", 1e30),                             # absurd count (would wrap size_t arithmetic)
                   lambda d: d["accessors"][1].__setitem__("byteOffset", 1e18),
                   lambda d: d["bufferViews"][0].__setitem__("byteOffset", -8),
                   lambda d: d["bufferViews"][0].__setitem__("byteLength", 10**9),                     # view longer than its buffer
                   lambda d: d["bufferViews"][0].__setitem__("byteStride", 4),                         # stride smaller than the element
                   lambda d: d["accessors"][0].__setitem__("count", 4),                                # one element past the view
                   lambda d: d["nodes"][0].__setitem__("children", [5]),                               # child index out of range
                   lambda d: d["nodes"][0].__setitem__("children", [0]),                               # cycle
                   lambda d: d["nodes"][0].__setitem__("children", ["x"]),
                   lambda d: d["scenes"][0].__setitem__("nodes", [-1]),
                   lambda d: d["nodes"][0].__setitem__("mesh", 4)):
        bad = copy.deepcopy(good)
        mutate(bad)
        assert rc(bad) == -3, (L.eid_last_error(), bad)
    assert rc("[" * 5000 + "]" * 5000, "deep.gltf") == -3 and b"nesting" in L.eid_last_error()
    assert rc(good) == 0                                                                               # the scene object is still usable


def test_gltf_without_scenes_uses_the_parentless_nodes_as_roots(tmp_path):
    """No `scenes`: only nodes that are nobody's child are roots; children are reached through their parents (once, with the parent's
    transform) — the hierarchy is not instanced a second time without it."""
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes()
    doc = {"asset": {"version": "2.0"},
           "nodes": [{"mesh": 0, "translation": [0, 0, 1]}, {"children": [0, 2], "translation": [10, 0, 0]}, {"mesh": 0}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}}]}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}],
           "buffers": [{"byteLength": 36, "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    p = tmp_path / "noscene.gltf"
    p.write_text(json.dumps(doc))
    s = eid.Scene(device=-1)
    s.load(str(p))
    info = s.info()
    assert info.nodeCount == 2 and info.triangleInstances == 2
    assert np.allclose(info.bboxMin[:], (10, 0, 0)) and np.allclose(info.bboxMax[:], (11, 1, 1))


def test_pipeline_layout_host_logic():
    """eid_group_pipeline_layout (csrc/pipeline.cu): ranks per stage, bands of multiples of 8 rows that cover the frame, padded height."""
    from eidola_b200 import abi
    expect = {2: (1, 0, 1), 3: (1, 1, 1), 4: (2, 1, 1), 5: (2, 2, 1), 8: (3, 3, 2)}
    for world, (nd, ni, np_) in expect.items():
        lays = [eid.Group.pipeline_layout(1080, world, r) for r in range(world)]
        assert all((l.nDirect, l.nIndirect, l.nPost) == (nd, ni, np_) for l in lays)
        assert len({l.paddedHeight for l in lays}) == 1 and lays[0].paddedHeight >= 1080 and lays[0].paddedHeight % 8 == 0
        for stage in (abi.STAGE_DIRECT, abi.STAGE_INDIRECT, abi.STAGE_POST):
            bands = sorted((l.y0, l.y1) for l in lays if l.stages & stage)
            assert bands, "every stage runs somewhere"
            assert bands[0][0] == 0 and bands[-1][1] >= 1080 and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(y0 % 8 == 0 and y1 % 8 == 0 and y1 <= lays[0].paddedHeight for y0, y1 in bands)
        # without indirect ranks the single post rank traces the indirect stage too
        if ni == 0:
            assert lays[-1].stages == abi.STAGE_INDIRECT | abi.STAGE_POST
    l = eid.Group.pipeline_layout(1080, 8, 3)
    assert (l.stages, l.index, l.count, l.y0, l.y1, l.paddedHeight) == (abi.STAGE_INDIRECT, 0, 3, 0, 360, 1088)
    l = eid.Group.pipeline_layout(144, 4, 3, (1, 1, 2))
    assert (l.stages, l.index, l.count, l.y0, l.y1) == (abi.STAGE_POST, 1, 2, 72, 144)
    for bad in [(1080, 1, 0, (0, 0, 0)), (1080, 4, 4, (0, 0, 0)), (1080, 4, 0, (1, 1, 1)), (1080, 4, 0, (2, 0, 2)), (0, 2, 0, (0, 0, 0)), (1080, 4, 0, (0, 2, 2))]:
        with pytest.raises(eid.EidolaError):
            eid.Group.pipeline_layout(bad[0], bad[1], bad[2], bad[3])
    a, b = eid.Group.random_id(), eid.Group.random_id()
    assert len(a) == 128 and a != b
