"""CPU tests that pin the oracle: known-answer vectors derived from the reference source,
the reference's own stand-alone-compilable files (oracle/_ref/libref.so, built from /root/reference when it
exists) and the committed outputs of those files (tests/golden/ref_vectors.npz) for machines without it."""
import ctypes as C
import os

import numpy as np
import pytest

import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import oracle_lib as ol  # noqa: E402
from make_golden import STRUCTS  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz")


def test_oracle_compiled_without_fma_contraction():
    assert ol.lib().orc_fp_contract_selftest() == 0


def test_tea_known_answers():     # random.glsl:34-48
    L = ol.lib()
    assert [L.orc_tea(a, b) for a, b in ((0, 0), (1, 0), (0, 1), (1037760, 12345))] == [0x741c187d, 0x8da6b311, 0x70d3aef1, 0x6230029f]


def test_pcg_rand_chain():        # random.glsl:59-65, 98-102
    L = ol.lib()
    st, v = np.zeros(4, np.uint32), np.zeros(4, np.float32)
    L.orc_rand_chain(L.orc_tea(0, 0), 4, st.ctypes.data, v.ctypes.data)
    assert list(st) == [0x46dfb666, 0x5477ab23, 0xa3758fc4, 0x92110c99]
    assert np.allclose(v, [0.560322881, 0.745306849, 0.335232019, 0.887807727], rtol=0, atol=1e-9)
    assert (v >= 0).all() and (v < 1).all()


def test_hash8bit():              # common.glsl:141-143
    L = ol.lib()
    assert [L.orc_hash8bit(x) for x in (0, 1, 255, 256, 257, 0x1234)] == [0x0, 0x01000000, 0xff000000, 0x01000000, 0x0, 0x26000000]


def test_struct_sizes_match_reference_header():
    want = dict(SceneCamera=336, VertexAttributes=32, GltfShadeMaterial=80, RtxState=100, InstanceData=24, LightSample=28,
                GISample=64, DirectReservoir=36, IndirectReservoir=76, ImptSampData=16, PuncLight=80, TrigLight=96,
                LightBufInfo=16, Tonemapper=48, SunAndSky=96)
    L = ol.lib()
    for k, v in want.items():
        assert L.orc_sizeof(k.encode()) == v, k
    z = np.load(GOLD)
    assert dict(zip(STRUCTS, z["sizes"].tolist())) == want          # sizes the reference's own header compiled to
    R = ol.ref()
    if R is not None:
        for k, v in want.items():
            assert R.ref_sizeof(k.encode()) == v, k


def _oracle_codec(z):
    L = ol.lib()
    enc = np.array([L.orc_compress_unit_vec(float(a), float(b), float(c)) for a, b, c in z["vec"]], np.uint32)
    dec = np.zeros((z["words"].size, 3), np.float32)
    tmp = np.zeros(3, np.float32)
    for i, w in enumerate(z["words"]):
        L.orc_decompress_unit_vec(int(w), tmp.ctypes.data)
        dec[i] = tmp
    packed = np.array([L.orc_pack_unorm4x8(np.ascontiguousarray(c).ctypes.data) for c in z["cols"]], np.uint32)
    return enc, dec, packed


def test_oct_codec_and_pack_match_reference_vectors():
    """compress_unit_vec / decompress_unit_vec / packUnorm4x8 (compress.glsl C++ branch) — bit-exact."""
    z = np.load(GOLD)
    enc, dec, packed = _oracle_codec(z)
    assert np.array_equal(enc, z["enc"])
    assert dec.tobytes() == z["dec"].tobytes()
    assert np.array_equal(packed, z["packed"])
    assert enc[6] == 0xb6da7fff                                   # (0, 0.6, 0.8), SURVEY.md §4
    assert np.allclose(dec[6], (0, 0.6, 0.8), atol=1e-4)


def test_alias_table_matches_reference_vectors():
    """DiscreteSampler1D<float> (alias_table.hpp:21-63) — bit-exact (prob, failId) per cell."""
    z = np.load(GOLD)
    L = ol.lib()
    off = 0
    for n in z["alias_n"]:
        w = np.ascontiguousarray(z["alias_in"][off:off + n])
        p, f = np.zeros(n, np.float32), np.zeros(n, np.int32)
        L.orc_alias_table(w.ctypes.data, int(n), p.ctypes.data, f.ctypes.data)
        assert p.tobytes() == z["alias_p"][off:off + n].tobytes() and np.array_equal(f, z["alias_f"][off:off + n])
        if n == 4:
            assert list(p) == [0.25, 0.5, 0.75, 1.0] and list(f) == [3, 3, 3, 3]      # {1,2,3,10}, SURVEY.md §4
        off += n


def test_shader_functions_match_reference_glsl_vectors():
    """The oracle's restatement of the reference's shader functions vs the reference's OWN GLSL text compiled as C++ (oracle/ref_shim:
    glsl_prep.py transliterates qualifiers / literals / swizzles / built-in names, the expressions stay the reference's): RNG, sampling
    helpers, the metallic-workflow BSDF (eval, pdf, sample), reservoir update / merge / validity / clamp, OffsetRay, the Uncharted-2
    tonemap and sun_and_sky() — bit for bit on seeded inputs (committed outputs: tests/golden/ref_vectors.npz)."""
    import ctypes as C
    import ref_fn_inputs as fi
    from eidola_b200 import abi
    z = np.load(GOLD)
    L = ol.lib()
    for w, (ni, no) in enumerate(fi.ARITY):
        got = ol.call_fn(L, "orc_fn", w, fi.inputs(w, n=1500), no)
        want = z["fn_%d_out" % w]
        bad = np.nonzero((got.view(np.uint32) != want.view(np.uint32)).any(axis=1))[0]
        assert bad.size == 0, "%s: %d of %d items differ from the reference GLSL" % (fi.NAMES[w], bad.size, len(got))
    dirs = np.ascontiguousarray(z["sky_dirs"])
    for k, kw in enumerate(fi.SKY_PARAMS):
        ss = fi.sun_sky(abi, kw)
        got = np.zeros_like(dirs)
        L.orc_sun_and_sky(C.byref(ss), dirs.ctypes.data, len(dirs), got.ctypes.data)
        assert got.view(np.uint32).tobytes() == z["sky_%d_out" % k].view(np.uint32).tobytes(), "sun_and_sky, parameter set %d" % k


def test_light_sampling_and_environment_match_reference_glsl_vectors():
    """pathtrace.glsl / env_sampling.glsl compiled as C++ against the oracle scene's own tables: SampleDirectLightNoVisibility (triangle
    lights, punctual lights, HDR alias map, sun & sky: pdf, Li, wi, dist and the RNG state after the draws), LightEval, EnvEval,
    EnvRadiance, raySpawn, clampRadiance, Sample — bit for bit on three scenes."""
    import common
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    z = np.load(GOLD)
    L = ol.lib()
    for tag, maker_name, kind in fi.CTX_CONFIGS:
        osc, orr, oenv, ss, st = ol.ctx_setup(scenes, abi, common, maker_name, kind)
        nmat = len(osc.table(abi.TABLE_MATERIALS))
        for w, (ni, no) in enumerate(fi.CTX_ARITY):
            key = "ctx_%s_%d_out" % (tag, w)
            if key not in z.files:
                continue
            x = np.ascontiguousarray(fi.ctx_inputs(w, nmat))
            got = np.zeros((x.shape[0], no), np.float32)
            assert L.orc_ctx_fn(orr._h, C.byref(st), w, x.ctypes.data, x.shape[0], got.ctypes.data) == 0
            bad = np.nonzero((got.view(np.uint32) != z[key].view(np.uint32)).any(axis=1))[0]
            assert bad.size == 0, "%s / %s: %d of %d items differ from the reference GLSL" % (tag, fi.CTX_NAMES[w], bad.size, len(got))


def _final_frame(renderer_factory, name):
    """Runs a golden config to its last frame with the given (scene, renderer) factory; returns the renderer and the last state."""
    import common
    import make_golden_cfg as cfg
    maker, size, frames, over = cfg.CONFIGS[name][:4]
    sc, rr = renderer_factory(maker(), size)
    sc.update_camera(*size)
    info = sc.info()
    for f in range(frames):
        sc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, **over)
        rr.run(st, f)
    return sc, rr, st, size


def test_post_stages_match_reference_shader_mains():
    """denoise_direct.comp x4, denoise_indirect.comp x5 and compose.comp — the reference's shader text INCLUDING main(), compiled as C++ and
    dispatched like Renderer::run (oracle/ref_shim/ref_post.cpp) — leave exactly the images the oracle's post stages leave
    (committed outputs: tests/golden/ref_post.npz; re-run live where /root/reference exists)."""
    import common
    from eidola_b200 import abi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_post.npz"))

    def factory(arrays, size):
        osc = ol.OracleScene()
        osc.load_arrays(arrays)
        orr = ol.OracleRenderer(osc, size)
        orr.set_env_constant(common.ENV)
        return osc, orr
    for name in ("c2_cornell", "room"):
        osc, orr, st, size = _final_frame(factory, name)
        for k in ("BUF_DIRECT", "BUF_INDIRECT", "BUF_DENOISE_IND_A", "BUF_DENOISE_IND_B"):
            assert orr.read(getattr(abi, k)).tobytes() == z["%s_%s" % (name, k)].tobytes(), (name, k)


def test_trace_stages_match_reference_shader_mains():
    """direct_stage.comp and indirect_stage.comp — the reference's shader text with everything it includes and its main(), compiled as C++
    and dispatched in 8x8 work groups like Renderer::run (oracle/ref_shim/ref_trace.cpp; the ray queries, which run inside the Vulkan
    driver, are answered by the oracle's intersector) — leave exactly the G-buffer, motion vectors, reservoirs and pre-denoise images the
    oracle leaves, frame after frame (temporal reuse included), with the same number of rays.  Committed outputs:
    tests/golden/ref_trace.npz; re-run live where /root/reference exists (test_live_reference_library_...)."""
    import common
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_trace.npz"))
    for c in fi.TRACE_CONFIGS:
        tag, maker_name, size, frames, kind, _ = c
        arrays, osc, orr, env, ss, over = ol.trace_setup(scenes, abi, common, c)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(*size)
            st = common.frame_state(size[0], size[1], info, f, **over)
            orr.run_trace(st, f, 0, size[1])
            if f < frames - 1:
                orr.run_post(st, f)
        got = ol.trace_snapshot(abi, orr)
        for k in fi.TRACE_KEYS:
            assert np.ascontiguousarray(got[k]).view(np.uint8).reshape(-1).tobytes() == z["%s_%s" % (tag, k)].tobytes(), (tag, k)
        s = orr.stats()
        assert [s.closestHitRays, s.anyHitRays] == [int(v) for v in z["%s_rays" % tag]], tag


def test_compile_time_variants_match_reference_shader_text():
    """The reference's dormant compile-time variants — DENOISER_DIRECT_BILATERAL (direct_stage.comp:284-288, denoise_direct.comp:73-137,
    renderer.cpp:186-188), DENOISER_INDIRECT_BILATERAL (denoise_indirect.comp:77-130) and FETCH_GEOM_CHECK_4_SUBPIXELS (pathtrace.glsl:314-358)
    — compiled from the reference's OWN shader text with the switches flipped (oracle/ref_shim, -DREF_VARIANT) and replayed frame by frame:
    the oracle with the same variant bits leaves exactly those buffers (committed: tests/golden/ref_variants.npz; live where /root/reference exists)."""
    import common
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_variants.npz"))
    for c in fi.VARIANT_CONFIGS:
        tag, maker_name, size, frames, variant, over = c
        arrays, osc, orr, env, ss, over = ol.trace_setup(scenes, abi, common, (tag, maker_name, size, frames, "none", over))
        orr.set_variant(variant)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(*size)
            orr.run(common.frame_state(size[0], size[1], info, f, **over), f)
        for k in fi.VARIANT_KEYS:
            got = np.ascontiguousarray(orr.read(getattr(abi, k))).view(np.uint8).reshape(-1)
            assert got.tobytes() == z["%s_%s" % (tag, k)].tobytes(), (tag, k)
    R = ol.ref()
    if R is not None:      # live: every buffer of EVERY frame, reference text vs oracle, and the committed vectors are what it produces
        sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
        import make_golden as mg
        for c in fi.VARIANT_CONFIGS:
            for k, v in mg.variant_replay(R, abi, scenes, c, check=True).items():
                assert v.tobytes() == z["%s_%s" % (c[0], k)].tobytes(), (c[0], k)


def test_oracle_display_pass_matches_reference_post_frag():
    """The oracle's restatement of RenderOutput::run / post.frag against the committed output of the reference's OWN post.frag (main()
    included, compiled as C++ by oracle/ref_shim/ref_display.cpp): every view and tonemapper of DISPLAY_CONFIGS, auto exposure included,
    bit for bit.  Committed outputs: tests/golden/ref_display.npz; re-run live where /root/reference exists."""
    import common
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_display.npz"))
    R = ol.ref() if os.path.isdir("/root/reference") else None
    frames = {}
    for tag, mode, over in fi.DISPLAY_CONFIGS:
        if mode not in frames:
            frames[mode] = ol.display_frames(scenes, abi, common, mode)
        orr, st, osc = frames[mode]
        tm = abi.default_tonemapper(**over)
        got = orr.run_output(tm, st)
        assert got.view(np.uint32).tobytes() == z["%s_out" % tag].view(np.uint32).tobytes(), tag
        w, h = fi.DISPLAY_SIZE
        d, i = orr.read(abi.BUF_DIRECT).reshape(h, w, 4), orr.read(abi.BUF_INDIRECT).reshape(h, w, 4)
        assert np.stack([ol.mip_chain_average(d), ol.mip_chain_average(i)]).tobytes() == z["%s_mips" % tag].tobytes(), tag
        if R is not None:
            assert ol.ref_display_run(R, tm, mode, d, i).view(np.uint32).tobytes() == z["%s_out" % tag].view(np.uint32).tobytes(), tag
    assert not np.array_equal(z["auto_out"], z["default_out"])


def test_mip_chain_average_known_answers():
    """The contract's restatement of nvvk::cmdGenerateMipmaps (linear blits down to 1x1): exact halvings are box filters, an odd extent
    samples between its texels at (i + 0.5) * src / dst - 0.5, a 1-wide axis is kept."""
    rng = np.random.default_rng(11)
    c = np.tile(np.array([0.25, 1.5, 3.0, 1.0], np.float32), (16, 32, 1))
    assert ol.mip_chain_average(c).tolist() == [0.25, 1.5, 3.0, 1.0]
    a = rng.random((2, 2, 4)).astype(np.float32)
    top, bot = a[0, 0] * np.float32(0.5) + a[0, 1] * np.float32(0.5), a[1, 0] * np.float32(0.5) + a[1, 1] * np.float32(0.5)
    assert ol.mip_chain_average(a).tobytes() == (top * np.float32(0.5) + bot * np.float32(0.5)).astype(np.float32).tobytes()
    a = rng.random((1, 3, 4)).astype(np.float32)          # 3 -> 1: u = 0.5 * 3 - 0.5 = 1.0, the middle texel
    assert ol.mip_chain_average(a).tobytes() == a[0, 1].tobytes()
    a = rng.random((8, 8, 4)).astype(np.float32)          # power of two: the plain mean up to rounding
    assert np.allclose(ol.mip_chain_average(a), a.reshape(-1, 4).mean(axis=0), rtol=1e-6)
    a = rng.random((5, 7, 4)).astype(np.float32)
    m = ol.mip_chain_average(a)
    assert (m >= a.reshape(-1, 4).min(axis=0)).all() and (m <= a.reshape(-1, 4).max(axis=0)).all()


def _zero_sign_free(punc_bytes, abi):
    """PuncLight table with the sign of zero direction / position components cleared: `worldMatrix * vec4(0, 0, -1, 0)` runs through
    nvmath's operator* (un-vendored) — whether a zero comes out as +0 or -0 is the stand-in's choice, not the reference's."""
    a = np.frombuffer(bytes(punc_bytes), abi.PUNC_DT).copy()
    for f in ("direction", "position"):
        v = a[f]
        v[v == 0.0] = 0.0
        a[f] = v
    return a.tobytes()


def test_host_tables_match_reference_scene_cpp():
    """Scene::load's table builders and Scene::updateCamera — the reference's OWN src/scene.cpp, compiled where it lies against stand-ins
    for Vulkan / nvpro_core / tinygltf (oracle/ref_shim/scene/) and run on the injected harness scene — against the oracle's restatement
    AND the product's host side: materials, punctual and triangle lights with their alias maps, LightBufInfo, every vertex and index
    buffer, instance material indices, both light weights, and the SceneCamera of a three-step camera sequence, bit for bit.  Committed
    outputs: tests/golden/ref_scene.npz; re-run live where /root/reference exists."""
    import ref_fn_inputs as fi
    import eidola_b200 as eid
    from eidola_b200 import abi, scenes
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_scene.npz"))
    live = ol.ref_scene_lib() if os.path.isdir("/root/reference") else None
    for name in fi.SCENE_TABLE_MAKERS:
        arrays = getattr(scenes, name)()
        osc = ol.OracleScene(); osc.load_arrays(arrays)
        psc = eid.Scene(device=-1); psc.load_arrays(arrays)          # host-only product scene: the table builders run on the CPU
        sides = [("oracle", osc), ("product", psc)]
        if live is not None:
            sides.append(("reference (live)", ol.RefScene(arrays)))
        for tag, side in sides:
            got = mg.scene_tables(abi, side, len(arrays.prim_meshes))
            for k, v in got.items():
                want = z["%s_%s" % (name, k)]
                if k == "punc":
                    assert _zero_sign_free(v, abi) == _zero_sign_free(want, abi), (name, tag, k)
                else:
                    assert v.tobytes() == want.tobytes(), (name, tag, k)
            inst = np.ascontiguousarray(side.table(abi.TABLE_INSTANCE_DATA)).view(np.uint8).reshape(-1).view(abi.INSTANCE_DT)["materialIndex"]
            assert np.array_equal(inst, z["%s_inst_material" % name]), (name, tag)
            w = side.weights() if hasattr(side, "weights") else (side.info().trigLightWeight, side.info().puncLightWeight)
            assert np.array(w, np.float32).tobytes() == z["%s_weights" % name].tobytes(), (name, tag)
            for k, (size, look) in enumerate(fi.SCENE_CAMERA_STEPS):
                if look is not None:
                    side.set_lookat(*look)
                side.update_camera(*size)
                cam = np.ascontiguousarray(side.table(abi.TABLE_CAMERA)).view(np.uint8).reshape(-1).view(np.uint32).copy()
                want = z["%s_camera_%d" % (name, k)].view(np.uint32).copy()
                if k == 0:
                    cam[80:83] = 0; want[80:83] = 0       # lastPosition of the first update: a function-static in the reference (scene.cpp:780), i.e. the previous scene's eye
                assert cam.tobytes() == want.tobytes(), (name, tag, "camera step %d" % k)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_instances_match_reference_accelstruct_cpp():
    """The reference's OWN src/accelstruct.cpp (compiled where it lies; nvvk::RaytracingBuilderKHR stands in and keeps what it is asked to
    build), called like SampleExample::loadScene does: per node the instance record — instanceCustomIndex = prim mesh, mask 0xFF, flags by
    the FORCE_OPAQUE / TRIANGLE_FACING_CULL_DISABLE rule of accelstruct.cpp:145-149, the node's world matrix as a 3x4 transform — and per prim
    mesh the geometry (indexCount / 3 triangles, 32-byte vertex stride, NO_DUPLICATE_ANY_HIT): what the oracle's intersector walks."""
    import ctypes as C
    import ref_fn_inputs as fi
    from eidola_b200 import scenes
    assert ol.ref_scene_lib() is not None
    for name in fi.SCENE_TABLE_MAKERS:
        arrays = getattr(scenes, name)()
        inst, xf, blas, build = ol.RefScene(arrays).accel()
        osc = ol.OracleScene(); osc.load_arrays(arrays)
        n = len(arrays.nodes)
        assert len(inst) == n and len(blas) == len(arrays.prim_meshes)
        flags = np.zeros((n, 3), np.int32)
        assert ol.lib().orc_scene_instance_flags(osc._h, flags.ctypes.data_as(C.c_void_p), n) == n
        oxf = np.zeros((n, 24), np.float32)
        assert ol.lib().orc_scene_instance_xforms(osc._h, oxf.ctypes.data_as(C.c_void_p), n) == n
        assert build == (4 | 2, 4)                                  # PREFER_FAST_TRACE | ALLOW_COMPACTION, PREFER_FAST_TRACE
        for i in range(n):
            custom, mask, sbt, fl, blas_idx = (int(v) for v in inst[i])
            assert (custom, mask, sbt, blas_idx) == (arrays.nodes[i]["primMesh"], 0xFF, 0, arrays.nodes[i]["primMesh"]), (name, i)
            assert (fl, custom) == (int(flags[i, 0]), int(flags[i, 1])), (name, i)
            assert int(blas[blas_idx, 0]) == int(flags[i, 2]), (name, i)                                # triangles of the instance
            o2w = oxf[i, :12].reshape(4, 3)                                                             # columns of objectToWorld
            assert xf[i].T.tobytes() == o2w.tobytes(), (name, i)                                        # row-major 3x4 == the same columns
        for k, pm in enumerate(arrays.prim_meshes):
            assert tuple(int(v) for v in blas[k]) == (pm["indexCount"] // 3, pm["vertexCount"], 32, 2, 106, 1), (name, k)   # R32G32B32_SFLOAT, UINT32
    assert any(int(f) != 4 for f in flags[:, 0])                    # the last scene (alpha) really has non-opaque instances


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_default_states_match_reference_initialisers(tmp_path):
    """The defaults every test, the bench and a drop-in host start from — abi.default_rtx_state / default_sun_and_sky / default_tonemapper and the
    C++ mirror's eidola::defaultRtxState — against the reference's OWN initialisers of SampleExample::m_rtxState, m_sunAndSky
    (sample_example.hpp:154-203) and RenderOutput::m_tm / m_depthTm (render_output.hpp:44-60), lifted from the headers and compiled."""
    import shutil
    import subprocess
    from eidola_b200 import abi
    assert ol.ref_scene_lib() is not None
    assert bytes(abi.default_rtx_state(0, 0)) == ol.ref_default_state(0)
    assert bytes(abi.default_sun_and_sky(in_use=0)) == ol.ref_default_state(1)
    assert bytes(abi.default_tonemapper()) == ol.ref_default_state(2)
    depth = np.frombuffer(ol.ref_default_state(3), np.float32)
    assert depth[:3].tolist() == [0.0, np.float32(2.2), 0.0] and not depth[3:].any()          # what test_display_pass_post_frag uses for eDepth
    # the fields SampleExample derives per scene / environment (sample_example.cpp:87, 104-105; statements lifted and compiled) against the
    # formulas of the test harness (tests/common.py) and of bench.py
    import common
    import ctypes as C
    for trig, punc, integral in ((3.25, 0.0, 3.1415927), (0.0, 125.664, 0.731), (17.5, 2.125, 11.0)):
        out = (C.c_float * 3)()
        ol.ref_scene_lib().ref_glue_state(trig, punc, integral, out)
        info = type("I", (), dict(trigLightWeight=np.float32(trig), puncLightWeight=np.float32(punc)))
        st = common.frame_state(8, 8, info, 0, **common.env_state_overrides(np.float32(integral)))
        assert (np.float32(st.lightLuminIntegInv), np.float32(st.fireflyClampThreshold), np.float32(st.envMapLuminIntegInv)) == (np.float32(out[0]), np.float32(out[1]), np.float32(out[2]))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "cis-565-final-vr-raytracer_b200")
    src = tmp_path / "d.cpp"
    src.write_text('#include <cstdio>\n#include "eidola.hpp"\nint main() { RtxState s = eidola::defaultRtxState(0, 0); fwrite(&s, sizeof s, 1, stdout); eidola::RenderOutput o; fwrite(&o.m_tm, sizeof o.m_tm, 1, stdout); fwrite(&o.m_depthTm, sizeof o.m_depthTm, 1, stdout); return 0; }\n')
    exe = str(tmp_path / "d")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    subprocess.check_call([cxx, "-std=c++17", "-I", os.path.join(root, "include"), "-I", os.path.join(pkg, "host"), str(src), "-L", pkg, "-leidola", "-o", exe])
    out = subprocess.run([exe], env=dict(os.environ, LD_LIBRARY_PATH=pkg), capture_output=True).stdout
    assert out == ol.ref_default_state(0) + ol.ref_default_state(2) + ol.ref_default_state(3)      # defaultRtxState, RenderOutput::m_tm, m_depthTm of the C++ mirror


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_reference_renderer_replay_matches_oracle_frames():
    """The reference end to end on the CPU: its OWN src/renderer.cpp decides what a frame executes — which descriptor set, which push
    constants, which pipeline, how many work groups (recorded by the stand-in device, ref_renderer.cpp) — and every recorded dispatch is
    replayed with the reference's OWN shader text (ref_trace.cpp / ref_post.cpp) on the resources its descriptor sets name.  Frame after
    frame, every buffer equals the oracle's (whose Renderer::run restates the same schedule by hand): G-buffers of both parities, motion,
    reservoirs, tempDirectResv, denoise temporaries, both result images."""
    import ctypes as C
    import common
    from eidola_b200 import abi, scenes
    R = ol.ref()
    assert R is not None and ol.ref_scene_lib() is not None
    configs = [("cornell_scene", (64, 40), 3, dict(maxDepth=3)), ("small_room", (50, 34), 3, dict(maxDepth=2, ReSTIRState=abi.eSpatiotemporal)),
               ("cube_scene", (48, 32), 2, dict(maxDepth=2, denoise=0))]
    configs += [(seed, (48, 32), 2, dict(maxDepth=2, RISSampleNum=3)) for seed in (1, 2, 3)]      # rooms lit by random directional / spot / point lights
    for maker_name, size, frames, over in configs:
        w, h = size
        arrays = _lit_scene(maker_name) if isinstance(maker_name, int) else getattr(scenes, maker_name)()
        osc = ol.OracleScene(); osc.load_arrays(arrays)
        orr = ol.OracleRenderer(osc, size); orr.set_env_constant((0.0, 0.0, 0.0))
        rt = ol.RefTracer(R, abi, arrays, osc, size)                   # scene tables + intersector binding of the trace stages
        rr = ol.RefRenderer(w, h)
        res = rr.resources()
        store = {}                                                      # resource id -> array, zero-initialised like the contract's history
        for role, (rid, kind, nbytes, _, _, _) in res.items():
            store[rid] = np.zeros(nbytes, np.uint8)
        direct, indirect = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)     # RenderOutput's result images (S_OUT)
        wiring = rr.wiring()
        osc.update_camera(w, h)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(w, h)
            st = common.frame_state(w, h, info, f, **over)
            orr.run(st, f)
            rows, pushes = rr.run(st, f)
            cur_set = cur_push = cur_pipe = None
            for what, a, b, c in rows:
                if what == 1: cur_set = c
                elif what == 2: cur_push = abi.RtxState.from_buffer_copy(pushes[c])
                elif what == 3: cur_pipe = a
                elif what == 4:
                    bound = lambda binding: store[wiring[(cur_set, binding)][0]]
                    if cur_pipe in (1, 4):                              # direct_stage / indirect_stage
                        assert (a, b) == ((cur_push.size.x // (1 if cur_pipe == 1 else 2) + 7) // 8, (cur_push.size.y // (1 if cur_pipe == 1 else 2) + 7) // 8)
                        rt.run(cur_push, f, direct=cur_pipe == 1, indirect=cur_pipe == 4,
                               bufs=dict(lastG=bound(0), thisG=bound(1), lastDR=bound(2), thisDR=bound(3), tempDR=bound(4), lastIR=bound(5), thisIR=bound(6),
                                         motion=bound(8), direct=direct, indirect=indirect, indA=bound(11)))
                    else:                                               # denoise_direct / denoise_indirect / compose
                        cam = np.ascontiguousarray(osc.table(abi.TABLE_CAMERA))
                        assert R.ref_post_dispatch(cur_pipe, C.addressof(cur_push), cam.ctypes.data, w, h, a, b, bound(1).ctypes.data, direct.ctypes.data,
                                                   indirect.ctypes.data, bound(9).ctypes.data, bound(10).ctypes.data, bound(11).ctypes.data, bound(12).ctypes.data) == 0
            s = (f + 1) % 2 + 1                                         # the set the reference bound this frame
            for which, binding in ((abi.BUF_THIS_GBUFFER, 1), (abi.BUF_LAST_GBUFFER, 0), (abi.BUF_MOTION, 8), (abi.BUF_THIS_DIRECT_RESV, 3), (abi.BUF_LAST_DIRECT_RESV, 2),
                                   (abi.BUF_THIS_INDIRECT_RESV, 6), (abi.BUF_LAST_INDIRECT_RESV, 5), (abi.BUF_TEMP_DIRECT_RESV, 4), (abi.BUF_DENOISE_DIR_A, 9),
                                   (abi.BUF_DENOISE_DIR_B, 10), (abi.BUF_DENOISE_IND_A, 11), (abi.BUF_DENOISE_IND_B, 12)):
                got = store[wiring[(s, binding)][0]]
                assert got.tobytes() == np.ascontiguousarray(orr.read(which)).view(np.uint8).reshape(-1).tobytes(), (maker_name, f, which)
            assert direct.tobytes() == orr.read(abi.BUF_DIRECT).tobytes() and indirect.tobytes() == orr.read(abi.BUF_INDIRECT).tobytes(), (maker_name, f)
            assert direct[..., :3].max() > 0.0 and store[wiring[(s, 3)][0]].any() and store[wiring[(s, 6)][0]].any()      # (the frames are not trivially empty)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_display_pass_host_side_matches_reference_render_output_cpp():
    """The reference's OWN src/render_output.cpp on the recording stand-in device: four RGBA32F result images with a full mip chain, the two
    S_OUT descriptor sets wired this = [i] / last = [!i] with the samplers on this, RenderOutput::run = push (m_tm — or m_depthTm in the depth
    view — with zoom and renderingRatio overwritten, then debugging_mode), post pipeline, set (frames + 1) % 2, one 3-vertex draw, and
    genMipmap = one chain per result image.  The recorded push constants, replayed through the reference's OWN post.frag on the oracle's result
    images, give the oracle's display image (with the tonemapper the reference itself selected)."""
    import ctypes as C
    import common
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    R = ol.ref()
    assert R is not None and ol.ref_scene_lib() is not None
    w, h = fi.DISPLAY_SIZE
    ro = ol.RefOutput(w, h)
    ids, mips = ro.roles()
    levels = int(np.floor(np.log2(max(w, h)))) + 1
    assert len(set(ids)) == 4 and mips == (levels,) * 4
    wiring = ro.wiring()
    for i in (0, 1):                                                    # OutputBindings: 0 eDirectSampler, 1 eIndirectSampler, 2 / 3 eLast*, 4 / 5 eThis*
        assert [wiring[(i + 1, b)] for b in range(6)] == [ids[i], ids[2 + i], ids[1 - i], ids[3 - i], ids[i], ids[2 + i]]
    tm_size = C.sizeof(abi.Tonemapper)
    for mode in (abi.eNoDebug, abi.eDirectStage, abi.eDepth):
        orr, st, osc = ol.display_frames(scenes, abi, common, mode)
        d, i_ = orr.read(abi.BUF_DIRECT).reshape(h, w, 4), orr.read(abi.BUF_INDIRECT).reshape(h, w, 4)
        for frames in (0, 1):
            rows, push = ro.run(st, 1.0, (1.0, 1.0), frames, gen_mips=True)
            assert rows == [(5, ids[0], levels, 1), (5, ids[1], levels, 1), (5, ids[2], levels, 1), (5, ids[3], levels, 1),
                            (2, 0, tm_size + 4, 0), (3, 9, 0, 0), (1, 0, 1, (frames + 1) % 2 + 1), (6, 3, 1, 0)]
            tm = abi.Tonemapper.from_buffer_copy(push[:tm_size])
            assert int.from_bytes(push[tm_size:tm_size + 4], "little", signed=True) == mode
            base = ol.ref_default_state(3 if mode == abi.eDepth else 2)                       # m_depthTm in the depth view, else m_tm
            want = abi.Tonemapper.from_buffer_copy(base); want.zoom = 1.0; want.renderingRatio = abi.Vec2(1.0, 1.0)
            assert bytes(tm) == bytes(want)
            got = ol.ref_display_run(R, tm, mode, d, i_)                                      # the reference's post.frag with the reference's push constants
            assert common.same_bits_or_both_nan(got, orr.run_output(tm, st)), (mode, frames)


def _random_scene(seed):
    """A small random scene that exercises the corners of Scene::load's table builders: several meshes under random affine node
    transforms (some instanced twice, some mirrored), materials with out-of-range ior, weakly / strongly emissive ones on both sides of the
    1e-2 luminance threshold of createTrigLightBuffer, every alpha mode, and point / directional / spot lights with random cones."""
    from eidola_b200 import abi as _abi
    SceneArrays = _abi.SceneArrays
    rng = np.random.default_rng(seed)
    f32 = lambda a: np.asarray(a, np.float32)
    nmat = int(rng.integers(2, 7))
    mats = []
    for m in range(nmat):
        e = rng.choice([0.0, 0.004, 0.02, 3.0, 40.0]) * rng.random(3)
        mats.append(SceneArrays.material(base=tuple(rng.random(4)), metallic=float(rng.random()), roughness=float(rng.random()), emissive=tuple(float(x) for x in e),
                                         double_sided=int(rng.integers(0, 2)), ior=float(rng.choice([0.3, 1.0, 1.45, 2.7, 9.0])), transmission=float(rng.random()),
                                         alpha_mode=int(rng.integers(0, 3)), alpha_cutoff=float(rng.random())))
    pos, nrm, tan, uv, col, idx, prims, nodes = [], [], [], [], [], [], [], []
    nv = ni = 0
    for k in range(int(rng.integers(1, 5))):
        n = int(rng.integers(3, 9))
        p = rng.normal(size=(n, 3)); nn = rng.normal(size=(n, 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
        t = rng.normal(size=(n, 3)); t /= np.linalg.norm(t, axis=1, keepdims=True)
        tris = rng.integers(0, n, size=(int(rng.integers(1, 6)), 3))
        pos.append(p); nrm.append(nn); tan.append(np.concatenate([t, rng.choice([-1.0, 1.0], size=(n, 1))], axis=1)); uv.append(rng.random((n, 2)) * 3 - 1)
        col.append(rng.random((n, 4))); idx.append(tris.reshape(-1))
        prims.append(dict(firstIndex=ni, indexCount=int(tris.size), vertexOffset=nv, vertexCount=n, materialIndex=int(rng.integers(0, nmat))))
        for _ in range(int(rng.integers(1, 3))):                        # the mesh is instanced once or twice
            a = rng.normal(size=(3, 3)) * rng.choice([1.0, -1.0])       # random affine map, mirrored half of the time
            mtx = np.eye(4); mtx[:3, :3] = a; mtx[:3, 3] = rng.normal(size=3) * 4
            nodes.append(dict(worldMatrix=[float(v) for v in f32(mtx).T.reshape(-1)], primMesh=k))
        nv += n; ni += int(tris.size)
    lights = []
    for _ in range(int(rng.integers(0, 6))):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        mtx = np.eye(4); mtx[:3, :3] = q; mtx[:3, 3] = rng.normal(size=3) * 5
        inner = float(rng.random() * 0.6)
        lights.append(dict(worldMatrix=[float(v) for v in f32(mtx).T.reshape(-1)], type=int(rng.integers(0, 3)), color=tuple(float(v) for v in rng.random(3)),
                           intensity=float(rng.random() * 50), range=float(rng.random() * 10), innerConeAngle=inner, outerConeAngle=inner + float(rng.random() * 0.7)))
    cam = dict(eye=(0.0, 1.0, -6.0), center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), yfov=float(rng.random() + 0.4))
    return SceneArrays(np.concatenate(pos), np.concatenate(nrm), np.concatenate(tan), np.concatenate(uv), np.concatenate(col), np.concatenate(idx), prims, nodes, mats,
                       lights, cam, "random%d" % seed)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("workload", ["c3", "c5"])
def test_full_size_host_tables_match_reference_scene_cpp(workload):
    """BASELINE.json's full-size scenes — C3 (1 000 708 triangles, 1 000 emissive) and C5 (10 009 402 triangles, 10 000 emissive) — through the
    reference's OWN scene.cpp and through the product's host side: every vertex / index buffer, the materials, the 1 000- / 10 000-entry
    triangle-light table with its alias map, LightBufInfo and the light weight are bit-identical."""
    import eidola_b200 as eid
    from eidola_b200 import abi, scenes
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    assert ol.ref_scene_lib() is not None
    arrays = scenes.heightfield_room() if workload == "c3" else scenes.heightfield_room(quads=2236, n_light_quads=5000, light_seed=567)
    ref = ol.RefScene(arrays)
    psc = eid.Scene(device=-1); psc.load_arrays(arrays)
    want, got = mg.scene_tables(abi, ref, len(arrays.prim_meshes)), mg.scene_tables(abi, psc, len(arrays.prim_meshes))
    assert sorted(want) == sorted(got)
    for k in want:
        assert got[k].tobytes() == want[k].tobytes(), k
    li = np.frombuffer(want["info"].tobytes(), abi.LIGHTINFO_DT)[0]
    assert int(li["trigLightSize"]) == (1000 if workload == "c3" else 10000) and psc.info().triangleInstances == (1000708 if workload == "c3" else 10009402)
    assert np.array((psc.info().trigLightWeight, psc.info().puncLightWeight), np.float32).tobytes() == np.array(ref.weights(), np.float32).tobytes()


def _lit_scene(seed):
    """A visible room (floor, back wall, two boxes) with random materials, lit ONLY by random punctual lights of all three kinds (point,
    directional, spot with random cones / ranges) and one emissive quad: exercises SamplePuncLight's directional and spot branches."""
    from eidola_b200 import scenes as sc
    rng = np.random.default_rng(1000 + seed)
    b = sc._Builder()
    mats = [b.add_material(base=tuple(float(v) for v in 0.2 + 0.7 * rng.random(3)) + (1.0,), metallic=float(rng.random() < 0.3) * float(rng.random()),
                           roughness=float(0.2 + 0.8 * rng.random())) for _ in range(4)]
    lamp = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(6.0, 5.0, 4.0))
    b.add_quads(sc._box_quads((-3.0, 0.0, -3.0), (3.0, 3.0, 3.0), "yZx", inward=True), mats[0])
    b.add_quads(sc._box_quads((-1.6, 0.0, -0.4), (-0.4, 1.3, 0.9), "xXYzZ"), mats[1])
    b.add_quads(sc._box_quads((0.3, 0.0, -1.0), (1.5, 0.7, 0.2), "xXYzZ"), mats[2])
    b.add_quads([[(-0.4, 2.9, -0.4), (0.4, 2.9, -0.4), (0.4, 2.9, 0.4), (-0.4, 2.9, 0.4)]], lamp)
    for k in range(4):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        if k < 3:                                                        # aim the -Z axis (the light's direction) roughly downwards
            d = np.array([rng.normal() * 0.4, -1.0, rng.normal() * 0.4]); d /= np.linalg.norm(d)
            x = np.cross([0.0, 0.0, 1.0], d); x /= np.linalg.norm(x); y = np.cross(-d, x)
            q = np.stack([x, y, -d], axis=1)
        mtx = np.eye(4); mtx[:3, :3] = q; mtx[:3, 3] = (rng.uniform(-2, 2), rng.uniform(1.5, 2.8), rng.uniform(-2, 2))
        inner = float(rng.uniform(0.1, 0.5))
        b.lights.append(dict(worldMatrix=[float(v) for v in np.asarray(mtx, np.float32).T.reshape(-1)], type=[0, 2, 2, 1][k], color=tuple(float(v) for v in 0.5 + 0.5 * rng.random(3)),
                             intensity=float(rng.uniform(2, 30)), range=float(rng.uniform(0, 8)), innerConeAngle=inner, outerConeAngle=inner + float(rng.uniform(0.05, 0.6))))
    cam = dict(eye=(0.0, 1.6, -2.8), center=(0.0, 0.9, 0.5), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(70.0)))
    return b.build(cam, "lit%d" % seed)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_random_scenes_host_tables_and_instances_match_reference():
    """Randomised scenes through the reference's OWN scene.cpp / accelstruct.cpp, the oracle and the product's host side: every table, both
    light weights, the camera and the instance records agree bit for bit (only the sign of a zero light-direction / position component,
    which nvmath's un-vendored operator* decides, is left open)."""
    import ctypes as C
    import eidola_b200 as eid
    from eidola_b200 import abi
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    assert ol.ref_scene_lib() is not None
    seen_trig = seen_punc = 0
    for seed in range(24):
        arrays = _random_scene(seed)
        ref = ol.RefScene(arrays)
        osc = ol.OracleScene(); osc.load_arrays(arrays)
        psc = eid.Scene(device=-1); psc.load_arrays(arrays)
        want = mg.scene_tables(abi, ref, len(arrays.prim_meshes))
        for tag, side in (("oracle", osc), ("product", psc)):
            got = mg.scene_tables(abi, side, len(arrays.prim_meshes))
            for k in want:
                if k == "punc":
                    assert _zero_sign_free(got[k], abi) == _zero_sign_free(want[k], abi), (seed, tag, k)
                else:
                    assert got[k].tobytes() == want[k].tobytes(), (seed, tag, k)
            assert np.array((side.info().trigLightWeight, side.info().puncLightWeight), np.float32).tobytes() == np.array(ref.weights(), np.float32).tobytes(), (seed, tag)
        ref.update_camera(320, 200); osc.update_camera(320, 200); psc.update_camera(320, 200)
        ref.update_camera(320, 200); osc.update_camera(320, 200); psc.update_camera(320, 200)
        cam = ref.table(abi.TABLE_CAMERA).tobytes()
        assert np.ascontiguousarray(osc.table(abi.TABLE_CAMERA)).tobytes() == cam and np.ascontiguousarray(psc.table(abi.TABLE_CAMERA)).tobytes() == cam, seed
        inst, xf, blas, _ = ref.accel()
        n = len(arrays.nodes)
        flags = np.zeros((n, 3), np.int32)
        assert ol.lib().orc_scene_instance_flags(osc._h, flags.ctypes.data_as(C.c_void_p), n) == n
        oxf = np.zeros((n, 24), np.float32)
        ol.lib().orc_scene_instance_xforms(osc._h, oxf.ctypes.data_as(C.c_void_p), n)
        for i in range(n):
            assert (int(inst[i, 3]), int(inst[i, 0]), int(blas[inst[i, 4], 0])) == tuple(int(v) for v in flags[i]), (seed, i)
            assert xf[i].T.tobytes() == oxf[i, :12].tobytes(), (seed, i)
        li = np.frombuffer(want["info"].tobytes(), abi.LIGHTINFO_DT)[0]
        seen_trig += int(li["trigLightSize"] > 0); seen_punc += int(li["puncLightSize"] > 0)
    assert seen_trig >= 5 and seen_punc >= 5                            # the generator really produced both kinds of lights


def expected_run_commands(w, h, denoise, frames):
    """Renderer::run as the oracle (oracle_shaders.cpp Renderer::run / runPost) and the product (render.cu launchFrame, fillParams) implement
    it: descriptor set (frames + 1) % 2, the caller's RtxState pushed once, K1 over ceil(W/8) x ceil(H/8) groups, K2 over the (W/2) x (H/2)
    image, then — only when denoise > 0 — 4 direct and 5 indirect A-Trous levels, each with denoiseLevel = i pushed first, then compose.
    Rows are (what, a, b, c): 1 bind sets (first, count, set number), 2 push (offset, size, push index), 3 bind pipeline (shader tag), 4 dispatch."""
    cd = lambda x: (x + 7) // 8
    full, half = (cd(w), cd(h), 1), (cd(w // 2), cd(h // 2), 1)
    rows, levels = [(1, 0, 1, (frames + 1) % 2 + 1), (2, 0, 100, 0), (3, 1, 0, 0), (4,) + full, (3, 4, 0, 0), (4,) + half, (3, 5, 0, 0)], [None]
    for i in range(4 if denoise > 0 else 0):
        rows += [(2, 0, 100, len(levels)), (4,) + full]; levels.append(i)
    rows.append((3, 6, 0, 0))
    for i in range(5 if denoise > 0 else 0):
        rows += [(2, 0, 100, len(levels)), (4,) + half]; levels.append(i)
    rows += [(3, 7, 0, 0), (4,) + full]
    return rows, levels


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_renderer_schedule_and_wiring_match_reference_renderer_cpp():
    """The reference's OWN src/renderer.cpp, compiled where it lies against a recording stand-in for Vulkan (oracle/ref_shim/ref_renderer.cpp):
    Renderer::create / update allocate exactly the buffer set the oracle and the product allocate (sizes per pixel, full-resolution denoise
    temporaries, one tempDirectResv), the two descriptor sets are wired last* = [i], this* = [!i] (renderer.cpp:341-375), and
    Renderer::run records exactly the command sequence both implement (renderer.cpp:154-206), for even / odd sizes, denoise on / off and
    both frame parities; after a resize the wiring names the new resources."""
    from eidola_b200 import abi
    assert ol.ref_scene_lib() is not None
    LAST_THIS = {0: ("gbuffer", 0), 1: ("gbuffer", 1), 2: ("directResv", 0), 3: ("directResv", 1), 5: ("indirectResv", 0), 6: ("indirectResv", 1)}
    FIXED = {4: "directTemp", 7: "indirectTemp", 8: "motion", 9: "denoiseTemp0", 10: "denoiseTemp1", 11: "denoiseTemp2", 12: "denoiseTemp3"}
    rr = ol.RefRenderer(100, 60)
    for (w, h) in ((100, 60), (1920, 1080), (65, 33)):
        rr.update(w, h)
        res = rr.resources()
        n, ni = w * h, (w // 2) * (h // 2)
        for k in (0, 1):
            assert res["gbuffer%d" % k][1:] == (1, 16 * n, w, h, 107)                       # VK_FORMAT_R32G32B32A32_UINT
            assert res["directResv%d" % k][1:3] == (0, abi.DIRECT_RESV_DT.itemsize * n)
            assert res["indirectResv%d" % k][1:3] == (0, abi.INDIRECT_RESV_DT.itemsize * ni)
        assert res["directTemp"][1:3] == (0, 36 * n) and res["indirectTemp"][1:3] == (0, 76 * ni)
        assert res["motion"][1:] == (1, 4 * n, w, h, 82)                                     # VK_FORMAT_R16G16_SINT
        for k in range(4):
            assert res["denoiseTemp%d" % k][1:] == (1, 16 * n, w, h, 109)                    # RGBA32F, full resolution even for the indirect pair
        if n <= 100 * 60:          # the oracle's renderer allocates the same set, byte for byte
            from eidola_b200 import scenes
            osc = ol.OracleScene(); osc.load_arrays(scenes.cube_scene())
            orr = ol.OracleRenderer(osc, (w, h))
            for which, role in ((abi.BUF_THIS_GBUFFER, "gbuffer0"), (abi.BUF_LAST_GBUFFER, "gbuffer1"), (abi.BUF_MOTION, "motion"),
                                (abi.BUF_THIS_DIRECT_RESV, "directResv0"), (abi.BUF_LAST_DIRECT_RESV, "directResv1"),
                                (abi.BUF_THIS_INDIRECT_RESV, "indirectResv0"), (abi.BUF_LAST_INDIRECT_RESV, "indirectResv1"),
                                (abi.BUF_TEMP_DIRECT_RESV, "directTemp"), (abi.BUF_DENOISE_DIR_A, "denoiseTemp0"), (abi.BUF_DENOISE_DIR_B, "denoiseTemp1"),
                                (abi.BUF_DENOISE_IND_A, "denoiseTemp2"), (abi.BUF_DENOISE_IND_B, "denoiseTemp3")):
                assert orr.read(which).nbytes == res[role][2], role
        wiring = rr.wiring()
        assert len(wiring) == 26
        for i in (0, 1):                                                                     # descriptor set i = set number i + 1
            for binding, (name, which) in LAST_THIS.items():
                idx = i if which == 0 else 1 - i                                             # eLast* -> [i], eThis* -> [!i]
                assert wiring[(i + 1, binding)][0] == res["%s%d" % (name, idx)][0], (i, binding)
            for binding, name in FIXED.items():
                assert wiring[(i + 1, binding)][0] == res[name][0], (i, binding)
        for denoise in (1, 0):
            for frames in (0, 1, 2, 7):
                st = abi.default_rtx_state(w, h, denoise=denoise, maxDepth=3, time=1234 + frames)
                rows, pushes = rr.run(st, frames)
                want_rows, levels = expected_run_commands(w, h, denoise, frames)
                assert rows == want_rows, (w, h, denoise, frames)
                raw = bytes(st)
                assert pushes[0] == raw
                off = abi.RtxState.denoiseLevel.offset
                for p, lvl in zip(pushes[1:], levels[1:]):
                    assert p[:off] == raw[:off] and p[off + 4:] == raw[off + 4:] and int.from_bytes(p[off:off + 4], "little", signed=True) == lvl
                # compose runs with the push constants of the last denoise level still bound (no push before it): denoiseLevel 4 when denoise > 0


def test_environment_alias_map_matches_reference_vectors():
    """HdrSampling::createEnvironmentAccel / buildAliasmap (src/hdr_sampling.cpp:107-242, the reference's own code compiled where it
    lies) — alias, q, pdf, aliasPdf of every texel, the integral and the average: bit-exact in the oracle AND in the product's host side."""
    import eidola_b200 as eid
    z = np.load(GOLD)
    for tag in ("a", "b"):
        img = z["env_%s_img" % tag]
        want = z["env_%s_accel" % tag].tobytes()
        integral, average = (float(v) for v in z["env_%s_stats" % tag])
        o = ol.OracleEnv(img)
        assert o.accel().tobytes() == want and o.get_integral() == integral and o.get_average() == average
        p = eid.HdrSampling(-1)      # host-only environment: the table builders run on the CPU
        p.set_pixels(img)
        assert p.accel().tobytes() == want and p.get_integral() == integral and p.get_average() == average


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_live_reference_library_agrees_with_committed_vectors():
    """The committed ref_vectors.npz really is what the reference's code produces (re-run it live)."""
    R = ol.ref()
    assert R is not None
    z = np.load(GOLD)
    enc = np.array([R.ref_compress_unit_vec(float(a), float(b), float(c)) for a, b, c in z["vec"][:3000]], np.uint32)
    assert np.array_equal(enc, z["enc"][:3000])
    tmp = np.zeros(3, np.float32)
    for w, want in zip(z["words"][:3000], z["dec"][:3000]):
        R.ref_decompress_unit_vec(int(w), tmp.ctypes.data)
        assert tmp.tobytes() == want.tobytes()
    import ctypes as C
    from eidola_b200 import abi
    for tag in ("a", "b"):
        img = np.ascontiguousarray(z["env_%s_img" % tag])
        acc = np.zeros(img.shape[0] * img.shape[1], abi.IMPT_DT)
        integ, avg = C.c_float(), C.c_float()
        R.ref_env_accel(img.ctypes.data, img.shape[1], img.shape[0], acc.ctypes.data, C.byref(integ), C.byref(avg))
        assert acc.tobytes() == z["env_%s_accel" % tag].tobytes() and [integ.value, avg.value] == list(z["env_%s_stats" % tag])
    import common
    import ref_fn_inputs as fi
    from eidola_b200 import scenes
    for w, (ni, no) in enumerate(fi.ARITY):
        assert ol.call_fn(R, "ref_fn", w, fi.inputs(w, n=1500), no).view(np.uint32).tobytes() == z["fn_%d_out" % w].view(np.uint32).tobytes(), fi.NAMES[w]
    for tag, maker_name, kind in fi.CTX_CONFIGS:
        osc, orr, oenv, ss, st = ol.ctx_setup(scenes, abi, common, maker_name, kind)
        keep = ol.ref_scene_set(R, osc, oenv, ss, st, abi)
        nmat = len(osc.table(abi.TABLE_MATERIALS))
        for w, (ni, no) in enumerate(fi.CTX_ARITY):
            key = "ctx_%s_%d_out" % (tag, w)
            if key in z.files:
                assert ol.call_fn(R, "ref_ctx_fn", w, fi.ctx_inputs(w, nmat), no).view(np.uint32).tobytes() == z[key].view(np.uint32).tobytes(), (tag, w)
        del keep
    # whole frames: the reference's five stage shaders (mains included) against the oracle, every frame, live
    zt = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_trace.npz"))
    for c in fi.TRACE_CONFIGS:
        tag, maker_name, size, frames, kind, _ = c
        arrays, osc, orr, env, ss, over = ol.trace_setup(scenes, abi, common, c)
        rt = ol.RefTracer(R, abi, arrays, osc, size, env=env, sun_sky=ss)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(*size)
            st = common.frame_state(size[0], size[1], info, f, **over)
            orr.run_trace(st, f, 0, size[1])
            got, want = rt.run(st, f), ol.trace_snapshot(abi, orr)
            for k in fi.TRACE_KEYS:
                assert np.ascontiguousarray(got[k]).view(np.uint8).reshape(-1).tobytes() == np.ascontiguousarray(want[k]).view(np.uint8).reshape(-1).tobytes(), (tag, f, k)
            pre = {k: orr.read(getattr(abi, k)).copy() for k in ol.POST_BUFS}
            orr.run_post(st, f)
            post = ol.ref_post_run(R, abi, osc.table(abi.TABLE_CAMERA), st, size, pre)
            for k in ol.POST_BUFS:
                assert post[k].tobytes() == orr.read(getattr(abi, k)).tobytes(), (tag, f, k)
        for k in fi.TRACE_KEYS:
            assert np.ascontiguousarray(got[k]).view(np.uint8).reshape(-1).tobytes() == zt["%s_%s" % (tag, k)].tobytes(), (tag, k)


def test_offset_ray_properties():  # common.glsl:98-113
    L = ol.lib()
    rng = np.random.default_rng(3)
    out = np.zeros(3, np.float32)
    for _ in range(200):
        p = (rng.normal(size=3) * 10).astype(np.float32)
        n = rng.normal(size=3)
        n = (n / np.linalg.norm(n)).astype(np.float32)
        L.orc_offset_ray(p.ctypes.data, n.ctypes.data, out.ctypes.data)
        assert np.dot(out.astype(np.float64) - p, n) > 0            # moved to the normal's side
        assert np.abs(out - p).max() < 1e-2 * max(1.0, np.abs(p).max())


def test_detmath_accuracy():
    """The deterministic libm stand-ins are accurate enough to be drop-ins for the GLSL built-ins."""
    L = ol.lib()
    x = np.linspace(-20, 20, 200001).astype(np.float32)
    y = np.zeros_like(x)
    out = np.zeros_like(x)
    for op, fn, tol in ((0, np.sin, 2e-7), (1, np.cos, 2e-7)):
        L.orc_detmath(op, x.ctypes.data, y.ctypes.data, x.size, out.ctypes.data)
        assert np.abs(out - fn(x.astype(np.float64))).max() < tol
    xe = np.linspace(-80, 80, 200001).astype(np.float32)
    L.orc_detmath(2, xe.ctypes.data, y.ctypes.data, xe.size, out.ctypes.data)
    assert np.abs(out / np.exp(xe.astype(np.float64)) - 1).max() < 2e-7
    xa = np.linspace(-1, 1, 100001).astype(np.float32)
    o2 = np.zeros_like(xa)
    L.orc_detmath(5, xa.ctypes.data, xa.ctypes.data, xa.size, o2.ctypes.data)
    assert np.abs(o2 - np.arcsin(xa.astype(np.float64))).max() < 5e-7
    L.orc_detmath(6, xa.ctypes.data, xa.ctypes.data, xa.size, o2.ctypes.data)
    assert np.abs(o2 - np.arccos(xa.astype(np.float64))).max() < 5e-7


def test_oracle_golden_frames_are_reproducible():
    """The committed frame dumps are what the oracle produces today (guards against silent oracle drift)."""
    import common
    import make_golden_cfg as cfg
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    from eidola_b200 import abi
    for name, entry in cfg.CONFIGS.items():
        maker, size, frames, over = entry[:4]
        z = np.load(os.path.join(gdir, "frames_%s.npz" % name))
        osc = ol.OracleScene()
        osc.load_arrays(maker())
        orr = ol.OracleRenderer(osc, size)
        orr.set_env_constant(common.ENV)
        ss = cfg.sun_sky_of(entry)
        if ss is not None:
            orr.set_sun_and_sky(ss)
        osc.update_camera(*size)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(*size)
            st = common.frame_state(size[0], size[1], info, f, **over)
            orr.run(st, f)
        for k, v in common.snapshot(orr).items():
            assert v.view(np.uint8).tobytes() == z[k].view(np.uint8).tobytes(), (name, k)
        disp = orr.run_output(abi.default_tonemapper(), st)
        assert disp.view(np.uint8).tobytes() == z["display"].view(np.uint8).tobytes(), (name, "display")
        assert np.isfinite(disp).all() and disp[..., :3].min() >= 0.0 and disp[..., :3].max() <= 1.0 + 1.0 / 255.0


def test_oracle_invariants():
    """Size-independent properties (SURVEY.md §8c iv): reservoir clamps, finite non-negative images, sky texels."""
    import common
    from eidola_b200 import abi, scenes
    size = (160, 96)
    osc = ol.OracleScene()
    osc.load_arrays(scenes.small_room())
    orr = ol.OracleRenderer(osc, size)
    orr.set_env_constant(common.ENV)
    osc.update_camera(*size)
    info = osc.info()
    for f in range(4):
        osc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, reservoirClamp=2)
        orr.run(st, f)
        s = common.snapshot(orr)
        assert s["direct_resv"]["num"].max() <= st.RISSampleNum * st.reservoirClamp
        assert s["indirect_resv"]["num"].max() <= 2 * st.reservoirClamp
        for k in ("direct", "indirect"):
            assert np.isfinite(s[k]).all() and (s[k] >= 0).all()
        assert (s["direct_resv"]["weight"] >= 0).all() and (s["indirect_resv"]["weight"] >= 0).all()
    st2 = orr.stats()
    n, ni = size[0] * size[1], (size[0] // 2) * (size[1] // 2)
    assert st2.closestHitRays <= n + ni * st.maxDepth and st2.anyHitRays <= n + ni * (st.maxDepth - 1)   # ray model of SURVEY §8(d)


def test_sun_and_sky_known_behaviour():
    """oracle/oracle_sunsky.cpp (shaders/sun_and_sky.glsl): no reference vectors exist for it, so pin its qualitative behaviour."""
    from eidola_b200 import abi
    L = ol.lib()

    def sky(ss, dirs):
        d = np.asarray(dirs, np.float32)
        d = np.ascontiguousarray(d / np.linalg.norm(d, axis=1, keepdims=True))
        out = np.zeros_like(d)
        L.orc_sun_and_sky(C.byref(ss), d.ctypes.data, len(d), out.ctypes.data)
        return out

    ss = abi.default_sun_and_sky(in_use=1)
    up, horizon, ground = sky(ss, [[0, 1, 0], [1, 0.02, 0], [0.3, -0.8, 0.1]])
    assert np.isfinite([up, horizon, ground]).all() and (up > 0).all()
    assert up[2] > up[0], "clear-sky zenith is blue"
    assert horizon.sum() > up.sum(), "Perez sky brightens towards the horizon"
    assert abs(ground[0] / ground[2] - 1.0) < 0.3, "below the horizon: grey ground colour times irradiance"
    sd = np.array([0.0, 0.78, 0.62], np.float64); sd /= np.linalg.norm(sd)
    off = sd + np.array([0.02, 0.0, 0.0])
    assert sky(ss, [off])[0].sum() > 100 * up.sum(), "sun disc dominates"
    assert (sky(abi.default_sun_and_sky(in_use=1, multiplier=0.0), [[0, 1, 0]]) == 0).all()          # sun_and_sky.glsl:475-478
    night = sky(abi.default_sun_and_sky(in_use=1, sun_direction=abi.Vec3(0.0, -1.0, 0.0)), [[0, 1, 0]])[0]
    assert np.allclose(night, np.float32(3.1415926535) * np.array([0.0, 0.0, 0.01], np.float32)), "midnight: night_color * pi"
