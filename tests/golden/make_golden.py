"""Regenerates tests/golden/*.npz.

  ref_*.npz     outputs of the REFERENCE's own code (oracle/_ref/libref.so = shaders/compress.glsl C++ branch,
                src/alias_table.hpp, shaders/host_device.h compiled from /root/reference) on seeded inputs;
                needs /root/reference, so the vectors are committed for machines that lack it.
  frames_*.npz  per-buffer dumps of the CPU oracle after N frames of the configs in tests/make_golden_cfg.py
                (the oracle itself is "parity unpinned" at whole-frame level: the reference cannot run here).

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import common  # noqa: E402
import make_golden_cfg as cfg  # noqa: E402
import oracle_lib as ol  # noqa: E402

STRUCTS = ["SceneCamera", "VertexAttributes", "GltfShadeMaterial", "RtxState", "InstanceData", "LightSample", "GISample",
           "DirectReservoir", "IndirectReservoir", "ImptSampData", "PuncLight", "TrigLight", "LightBufInfo", "Tonemapper",
           "SunAndSky"]


def ref_vectors():
    R = ol.ref()
    if R is None:
        print("oracle/_ref/libref.so unavailable (no /root/reference): keeping the committed ref_*.npz")
        return
    rng = np.random.default_rng(20221212)
    v = rng.normal(size=(20000, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [0, 0.6, 0.8], [0.5, 0.5, -0.70710678]], np.float32)
    v = np.concatenate([axes, v]).astype(np.float32)
    enc = np.array([R.ref_compress_unit_vec(float(a), float(b), float(c)) for a, b, c in v], np.uint32)
    words = np.concatenate([enc[:4000], rng.integers(0, 2**32 - 1, 4000, dtype=np.uint64).astype(np.uint32)])
    dec = np.zeros((words.size, 3), np.float32)
    tmp = np.zeros(3, np.float32)
    for i, w in enumerate(words):
        R.ref_decompress_unit_vec(int(w), tmp.ctypes.data)
        dec[i] = tmp
    cols = rng.random((4000, 4)).astype(np.float32) * 1.2 - 0.1
    packed = np.array([R.ref_pack_unorm4x8(np.ascontiguousarray(c).ctypes.data) for c in cols], np.uint32)
    alias_in, alias_p, alias_f = [], [], []
    for n in (1, 2, 4, 7, 33, 1000):
        w = (rng.random(n).astype(np.float32) ** 3 * 20 + 0.01).astype(np.float32)
        if n == 4:
            w = np.array([1, 2, 3, 10], np.float32)
        p, f = np.zeros(n, np.float32), np.zeros(n, np.int32)
        R.ref_alias_table(w.ctypes.data, n, p.ctypes.data, f.ctypes.data)
        alias_in.append(w)
        alias_p.append(p)
        alias_f.append(f)
    # HdrSampling::createEnvironmentAccel (src/hdr_sampling.cpp:173-242, compiled where it lies) on two synthetic skies
    import ctypes as C
    from eidola_b200 import abi, scenes
    env = {}
    for tag, (w, h, seed, sun) in (("a", (32, 16, 9, True)), ("b", (24, 10, 4, False))):
        img = np.ascontiguousarray(scenes.synthetic_sky(w, h, seed, sun), np.float32)
        acc = np.zeros(w * h, abi.IMPT_DT)
        integ, avg = C.c_float(), C.c_float()
        R.ref_env_accel(img.ctypes.data, w, h, acc.ctypes.data, C.byref(integ), C.byref(avg))
        env["env_%s_img" % tag] = img
        env["env_%s_accel" % tag] = acc.view(np.uint8)
        env["env_%s_stats" % tag] = np.array([integ.value, avg.value], np.float32)
    # the reference's own GLSL text (random / common / pbr_metallicworkflow / reservoir / tonemapping / sun_and_sky .glsl) compiled as
    # C++ by oracle/ref_shim (glsl_prep.py + ref_glsl.cpp) on the seeded inputs of tests/ref_fn_inputs.py
    import ref_fn_inputs as fi
    for w, (ni, no) in enumerate(fi.ARITY):
        env["fn_%d_out" % w] = ol.call_fn(R, "ref_fn", w, fi.inputs(w, n=1500), no)
    drng = np.random.default_rng(77)
    dirs = drng.normal(size=(3000, 3))
    dirs[:600] = np.array([0.0, 0.78, 0.62]) / 0.99639 + 0.06 * dirs[:600]
    dirs = np.ascontiguousarray((dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32))
    env["sky_dirs"] = dirs
    for k, kw in enumerate(fi.SKY_PARAMS):
        ss = fi.sun_sky(abi, kw)
        out = np.zeros_like(dirs)
        R.ref_sun_and_sky(C.byref(ss), dirs.ctypes.data, len(dirs), out.ctypes.data)
        env["sky_%d_out" % k] = out
    # scene-dependent functions of pathtrace.glsl / env_sampling.glsl (light sampling, environment, camera rays) on three scenes
    for tag, maker_name, kind in fi.CTX_CONFIGS:
        osc, orr, oenv, ss, st = ol.ctx_setup(scenes, abi, common, maker_name, kind)
        keep = ol.ref_scene_set(R, osc, oenv, ss, st, abi)
        nmat = len(osc.table(abi.TABLE_MATERIALS))
        for w, (ni, no) in enumerate(fi.CTX_ARITY):
            if kind == "none" and w in (2, 3):
                continue          # no environment bound in the reference build (the constant environment is this repo's extension)
            env["ctx_%s_%d_out" % (tag, w)] = ol.call_fn(R, "ref_ctx_fn", w, fi.ctx_inputs(w, nmat), no)
        del keep
    sizes = np.array([R.ref_sizeof(s.encode()) for s in STRUCTS], np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), vec=v, enc=enc, words=words, dec=dec, cols=cols, packed=packed,
                        alias_in=np.concatenate(alias_in), alias_p=np.concatenate(alias_p), alias_f=np.concatenate(alias_f),
                        alias_n=np.array([a.size for a in alias_in], np.int32), sizes=sizes, **env)
    print("wrote ref_vectors.npz")


def frame_dumps():
    from eidola_b200 import abi
    for name, entry in cfg.CONFIGS.items():
        maker, size, frames, over = entry[:4]
        arrays = maker()
        osc = ol.OracleScene()
        osc.load_arrays(arrays)
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
        snap = common.snapshot(orr)
        snap["display"] = orr.run_output(abi.default_tonemapper(), st)
        np.savez_compressed(os.path.join(HERE, "frames_%s.npz" % name), **{k: v.view(np.uint8) if v.dtype.fields else v for k, v in snap.items()})
        print("wrote frames_%s.npz" % name)


def ref_post_vectors():
    """tests/golden/ref_post.npz: the reference's post-stage shaders (mains included, compiled as C++) run on the oracle's pre-denoise
    buffers of the last frame of two golden configs."""
    from eidola_b200 import abi
    R = ol.ref()
    if R is None:
        print("oracle/_ref/libref.so unavailable: keeping the committed ref_post.npz")
        return
    out = {}
    for name in ("c2_cornell", "room"):
        maker, size, frames, over = cfg.CONFIGS[name][:4]
        osc = ol.OracleScene()
        osc.load_arrays(maker())
        orr = ol.OracleRenderer(osc, size)
        orr.set_env_constant(common.ENV)
        osc.update_camera(*size)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(*size)
            st = common.frame_state(size[0], size[1], info, f, **over)
            orr.run_trace(st, f, 0, size[1])
            if f == frames - 1:
                pre = {k: orr.read(getattr(abi, k)).copy() for k in ol.POST_BUFS}
                post = ol.ref_post_run(R, abi, osc.table(abi.TABLE_CAMERA), st, size, pre)
                for k in ("BUF_DIRECT", "BUF_INDIRECT", "BUF_DENOISE_IND_A", "BUF_DENOISE_IND_B"):
                    out["%s_%s" % (name, k)] = post[k]
                assert post["BUF_DIRECT"].tobytes() != pre["BUF_DIRECT"].tobytes()
            orr.run_post(st, f)
    np.savez_compressed(os.path.join(HERE, "ref_post.npz"), **out)
    print("wrote ref_post.npz")


def ref_trace_vectors():
    """tests/golden/ref_trace.npz: what the reference's direct_stage.comp / indirect_stage.comp (main() included, compiled as C++ by
    oracle/ref_shim/ref_trace.cpp, ray queries answered by the oracle's intersector) leave after the last frame of each
    tests/ref_fn_inputs.TRACE_CONFIGS entry: G-buffer, motion vectors, both reservoir buffers, the two pre-denoise images, ray counts."""
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    R = ol.ref()
    if R is None:
        print("oracle/_ref/libref.so unavailable: keeping the committed ref_trace.npz")
        return
    out = {}
    for c in fi.TRACE_CONFIGS:
        tag, maker_name, size, frames, kind, _ = c
        arrays, osc, orr, env, ss, over = ol.trace_setup(scenes, abi, common, c)
        rt = ol.RefTracer(R, abi, arrays, osc, size, env=env, sun_sky=ss)
        info = osc.info()
        for f in range(frames):
            osc.update_camera(*size)
            got = rt.run(common.frame_state(size[0], size[1], info, f, **over), f)
        for k in fi.TRACE_KEYS:
            out["%s_%s" % (tag, k)] = np.ascontiguousarray(got[k]).view(np.uint8).reshape(-1).copy()
        out["%s_rays" % tag] = rt.rays.copy()
    np.savez_compressed(os.path.join(HERE, "ref_trace.npz"), **out)
    print("wrote ref_trace.npz")


def variant_replay(R, abi, scenes, cfgv, check=True):
    """One VARIANT_CONFIGS entry: every frame is rendered by the oracle (variant switched on) AND by the reference's own shader text from the
    variant builds (trace stages, then the post schedule) on its own buffers; returns the reference's buffers after the last frame.  With
    `check`, every buffer of every frame must agree between the two."""
    tag, maker_name, size, frames, variant, over = cfgv
    arrays, osc, orr, env, ss, over = ol.trace_setup(scenes, abi, common, (tag, maker_name, size, frames, "none", over))   # (black constant environment)
    orr.set_variant(variant)
    rt = ol.RefTracer(R, abi, arrays, osc, size, variant=variant)
    osc.update_camera(*size)
    info = osc.info()
    w, h = size
    ref = {}
    dirB, indB = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    for f in range(frames):
        osc.update_camera(*size)
        st = common.frame_state(w, h, info, f, **over)
        orr.run(st, f)
        got = rt.run(st, f)
        pre = {"BUF_THIS_GBUFFER": got["gbuffer"], "BUF_DIRECT": rt.direct, "BUF_INDIRECT": rt.indirect, "BUF_DENOISE_DIR_A": rt.dirA,
               "BUF_DENOISE_DIR_B": dirB, "BUF_DENOISE_IND_A": rt.indA, "BUF_DENOISE_IND_B": indB}
        post = ol.ref_post_run_variant(R, abi, osc.table(abi.TABLE_CAMERA), st, size, pre, variant)
        rt.direct[...] = post["BUF_DIRECT"].reshape(rt.direct.shape); rt.indirect[...] = post["BUF_INDIRECT"].reshape(rt.indirect.shape)
        rt.dirA[...] = post["BUF_DENOISE_DIR_A"].reshape(rt.dirA.shape); rt.indA[...] = post["BUF_DENOISE_IND_A"].reshape(rt.indA.shape)
        dirB[...] = post["BUF_DENOISE_DIR_B"].reshape(dirB.shape); indB[...] = post["BUF_DENOISE_IND_B"].reshape(indB.shape)
        ref = {"BUF_THIS_GBUFFER": got["gbuffer"], "BUF_MOTION": got["motion"], "BUF_THIS_DIRECT_RESV": got["direct_resv"],
               "BUF_THIS_INDIRECT_RESV": got["indirect_resv"], "BUF_DIRECT": rt.direct, "BUF_INDIRECT": rt.indirect, "BUF_DENOISE_DIR_A": rt.dirA,
               "BUF_DENOISE_IND_A": rt.indA, "BUF_DENOISE_IND_B": indB}
        if check:
            for k, v in ref.items():
                a = np.ascontiguousarray(v).view(np.uint8).reshape(-1)
                b = np.ascontiguousarray(orr.read(getattr(abi, k))).view(np.uint8).reshape(-1)
                assert a.tobytes() == b.tobytes(), "variant %s frame %d: %s differs between the reference text and the oracle" % (tag, f, k)
            s = orr.stats()
            assert [s.closestHitRays, s.anyHitRays] == [int(v) for v in rt.rays], (tag, f)
    return {k: np.ascontiguousarray(v).view(np.uint8).reshape(-1).copy() for k, v in ref.items()}


def ref_variant_vectors():
    """tests/golden/ref_variants.npz: the reference's compile-time variants (bilateral denoisers, FETCH_GEOM_CHECK_4_SUBPIXELS) run from the
    variant builds of its own shader text (oracle/_ref, -DREF_VARIANT): every buffer after the last frame of each VARIANT_CONFIGS entry."""
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    R = ol.ref()
    if R is None:
        print("oracle/_ref/libref.so unavailable: keeping the committed ref_variants.npz")
        return
    out = {}
    for c in fi.VARIANT_CONFIGS:
        for k, v in variant_replay(R, abi, scenes, c).items():
            out["%s_%s" % (c[0], k)] = v
    np.savez_compressed(os.path.join(HERE, "ref_variants.npz"), **out)
    print("wrote ref_variants.npz")


def ref_display_vectors():
    """tests/golden/ref_display.npz: what the reference's post.frag (main() included, compiled as C++ by oracle/ref_shim/ref_display.cpp)
    makes of the oracle's result images for every tests/ref_fn_inputs.DISPLAY_CONFIGS entry, plus the two 1x1 mip texels it was handed."""
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    R = ol.ref()
    if R is None:
        print("oracle/_ref/libref.so unavailable: keeping the committed ref_display.npz")
        return
    out, frames = {}, {}
    for tag, mode, over in fi.DISPLAY_CONFIGS:
        if mode not in frames:
            orr, st, osc = ol.display_frames(scenes, abi, common, mode)
            w, h = fi.DISPLAY_SIZE
            frames[mode] = (orr.read(abi.BUF_DIRECT).reshape(h, w, 4).copy(), orr.read(abi.BUF_INDIRECT).reshape(h, w, 4).copy())
        d, i = frames[mode]
        out["%s_out" % tag] = ol.ref_display_run(R, abi.default_tonemapper(**over), mode, d, i)
        out["%s_mips" % tag] = np.stack([ol.mip_chain_average(d), ol.mip_chain_average(i)])
    np.savez_compressed(os.path.join(HERE, "ref_display.npz"), **out)
    print("wrote ref_display.npz")


def scene_tables(abi, side, n_prim):
    """Tables 0..6 of a loaded scene object (reference / oracle / product: same `table(which, index)` call) as raw bytes."""
    raw = lambda a: np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()
    out = {"materials": raw(side.table(abi.TABLE_MATERIALS)), "punc": raw(side.table(abi.TABLE_PUNC_LIGHTS)), "trig": raw(side.table(abi.TABLE_TRIG_LIGHTS)),
           "info": raw(side.table(abi.TABLE_LIGHT_INFO))}
    out["info"][12:16] = 0            # LightBufInfo.pad: never written by the reference (whatever the Scene object's memory held)
    if not out["info"][:8].any():
        out["info"][8:12] = 0         # no lights at all: trigSampProb is not assigned either (scene.cpp:101-103)
    out["materials"].reshape(-1, 80)[:, 76:80] = 0      # GltfShadeMaterial.pad: `GltfShadeMaterial smat;` is not initialised there (scene.cpp:429)
    for pm in range(n_prim):
        out["vertices_%d" % pm] = raw(side.table(abi.TABLE_VERTICES, pm))
        out["indices_%d" % pm] = raw(side.table(abi.TABLE_INDICES, pm))
    return out


def ref_scene_vectors():
    """tests/golden/ref_scene.npz: the tables the reference's OWN src/scene.cpp (compiled where it lies against the stand-ins of
    oracle/ref_shim/scene/, run on the injected harness scene) uploads — materials, punctual / triangle lights with their alias maps,
    LightBufInfo, per-prim-mesh vertex and index buffers, the instance material indices, both light weights — and the SceneCamera after
    each step of SCENE_CAMERA_STEPS."""
    import ref_fn_inputs as fi
    from eidola_b200 import abi, scenes
    if ol.ref_scene_lib() is None:
        print("oracle/_ref/librefscene.so unavailable: keeping the committed ref_scene.npz")
        return
    out = {}
    for name in fi.SCENE_TABLE_MAKERS:
        arrays = getattr(scenes, name)()
        r = ol.RefScene(arrays)
        for k, v in scene_tables(abi, r, len(arrays.prim_meshes)).items():
            out["%s_%s" % (name, k)] = v
        out["%s_inst_material" % name] = r.table(abi.TABLE_INSTANCE_DATA).view(abi.INSTANCE_DT)["materialIndex"].copy()
        out["%s_weights" % name] = np.array(r.weights(), np.float32)
        for k, (size, look) in enumerate(fi.SCENE_CAMERA_STEPS):
            if look is not None:
                r.set_lookat(*look)
            r.update_camera(*size)
            out["%s_camera_%d" % (name, k)] = r.table(abi.TABLE_CAMERA).copy()
    np.savez_compressed(os.path.join(HERE, "ref_scene.npz"), **out)
    print("wrote ref_scene.npz")


if __name__ == "__main__":
    ref_scene_vectors()
    ref_display_vectors()
    ref_vectors()
    ref_post_vectors()
    ref_trace_vectors()
    ref_variant_vectors()
    frame_dumps()
