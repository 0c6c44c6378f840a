"""Seeded inputs of the function taps (orc_fn / ref_fn): shared by tests/golden/make_golden.py and tests/test_oracle_kat.py."""
import numpy as np

NAMES = ["toConcentricDisk", "powerHeuristic", "GetSphericalUv", "CreateCoordinateSystem", "HDRToLDR", "LDRToHDR", "metallicWorkflowBSDF",
         "metallicWorkflowPdf", "metallicWorkflowSample", "DirectReservoir ops", "IndirectReservoir ops", "toneMap", "OffsetRay", "tea", "rand"]
ARITY = [(2, 2), (2, 1), (3, 2), (3, 6), (3, 3), (3, 3), (14, 3), (14, 1), (14, 7), (30, 27), (39, 19), (4, 3), (6, 3), (2, 1), (2, 3)]


def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _bits(u):
    return np.asarray(u, np.uint32).view(np.float32)


def inputs(which, n=4000, seed=2022):
    rng = np.random.default_rng(seed + which)
    f = lambda *shape: rng.random(shape).astype(np.float32)   # noqa: E731
    if which == 0:
        return f(n, 2)
    if which == 1:
        return (f(n, 2) * np.float32(50.0)).astype(np.float32)
    if which in (2, 3):
        v = _unit(rng, n)
        v[:6] = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [1, 0, 0], [0, 0.001, 0.9999995], [0.6, 0, 0.8]], np.float32)
        return v
    if which in (4, 5):
        return (f(n, 3) * np.float32(4.0) if which == 4 else f(n, 3)).astype(np.float32)
    if which in (6, 7, 8):
        albedo, rough, metal = f(n, 3), np.maximum(f(n, 1), np.float32(0.001)), f(n, 1)
        metal[: n // 4] = 0.0
        metal[n // 4: n // 2] = 1.0
        nrm = _unit(rng, n)
        wo = _unit(rng, n)
        wo = np.where((np.sum(wo * nrm, axis=1, keepdims=True) < 0) & (rng.random((n, 1)) < 0.8), -wo, wo).astype(np.float32)
        last = _unit(rng, n) if which != 8 else f(n, 3)
        return np.concatenate([albedo, rough, metal, nrm, wo, last], axis=1).astype(np.float32)
    if which == 9:
        x = (f(n, 30) * np.float32(3.0)).astype(np.float32)
        x[:, 7] = _bits(rng.integers(0, 400, n))
        x[:, 25] = _bits(rng.integers(0, 400, n))
        x[:, 17], x[:, 27] = f(n), f(n)
        x[:, 28] = rng.integers(1, 330, n).astype(np.float32)
        x[: n // 10, 8] = np.nan
        x[n // 10: n // 5, 26] = -1.0
        return x
    if which == 10:
        x = (f(n, 39) * np.float32(3.0)).astype(np.float32)
        x[:, 16] = _bits(rng.integers(0, 400, n))
        x[:, 36] = f(n)
        x[:, 37] = rng.integers(1, 170, n).astype(np.float32)
        x[: n // 10, 17] = np.nan
        return x
    if which == 11:
        x = (f(n, 4) * np.float32(6.0)).astype(np.float32)
        x[:, 3] = f(n) * np.float32(2.0) + np.float32(0.1)
        return x
    if which == 12:
        p = ((f(n, 3) - np.float32(0.5)) * np.float32(40.0)).astype(np.float32)
        p[: n // 4] *= np.float32(0.001)
        return np.concatenate([p, _unit(rng, n)], axis=1).astype(np.float32)
    if which in (13, 14):
        return _bits(rng.integers(0, 2**32 - 1, (n, 2), dtype=np.uint64).astype(np.uint32))
    raise ValueError(which)


SKY_PARAMS = [dict(), dict(haze=5.0, saturation=1.4, redblueshift=-0.2, horizon_height=-0.7, horizon_blur=0.0),
              dict(sun_direction=(0.7, -0.1, 0.2), physically_scaled_sun=0, y_is_up=0), dict(sun_direction=(0.2, -0.6, 0.3), haze=2.0)]


def sun_sky(abi, kw):
    kw = dict(kw, in_use=1)
    if "sun_direction" in kw:
        kw["sun_direction"] = abi.Vec3(*kw["sun_direction"])
    return abi.default_sun_and_sky(**kw)


# ---- scene-dependent functions (orc_ctx_fn / ref_ctx_fn / eid_renderer_fn_tap) ------------------------------------------------------
CTX_NAMES = ["SampleDirectLightNoVisibility", "LightEval", "EnvEval", "EnvRadiance", "raySpawn", "clampRadiance", "Sample"]
CTX_ARITY = [(4, 9), (9, 4), (3, 4), (3, 3), (4, 6), (3, 3), (12, 8)]
CTX_SIZE = (96, 64)
# (tag, scene maker name, environment kind): lights only / HDR map + point light / sun & sky + emissive triangles
CTX_CONFIGS = [("cornell_none", "cornell_scene", "none"), ("cube_hdr", "cube_scene", "hdr"), ("room_sky", "small_room", "sky")]
CTX_SKY = dict(sun_direction=(0.3, 0.5, 0.4), haze=1.0)


def ctx_env_image(scenes):
    return np.ascontiguousarray(scenes.synthetic_sky(32, 16, 9, True), np.float32)


def ctx_inputs(which, n_materials, n=1200, seed=99):
    rng = np.random.default_rng(seed + which)
    unit = lambda: _unit(rng, n)   # noqa: E731
    bits = rng.integers(0, 2**32 - 1, (n, 1), dtype=np.uint64).astype(np.uint32).view(np.float32)
    w, h = CTX_SIZE
    if which == 0:
        return np.concatenate([bits, ((rng.random((n, 3)) - 0.5) * 4).astype(np.float32)], axis=1)
    if which == 1:
        mat = rng.integers(0, n_materials, (n, 1)).astype(np.uint32).view(np.float32)
        return np.concatenate([mat, (rng.random((n, 1)) * 5 + 0.1).astype(np.float32), unit(), unit(), (rng.random((n, 1)) + 0.01).astype(np.float32)], axis=1)
    if which in (2, 3):
        return unit()
    if which == 4:
        return np.concatenate([rng.integers(0, w, (n, 1)), rng.integers(0, h, (n, 1)), np.full((n, 1), w), np.full((n, 1), h)], axis=1).astype(np.float32)
    if which == 5:
        x = (rng.random((n, 3)) * 30).astype(np.float32)
        x[:20, 1] = np.nan
        return x
    if which == 6:
        nrm = unit()
        v = unit()
        v = np.where(np.sum(v * nrm, axis=1, keepdims=True) < 0, -v, v).astype(np.float32)
        return np.concatenate([bits, rng.random((n, 3)).astype(np.float32), np.maximum(rng.random((n, 1)), 0.001).astype(np.float32),
                               rng.random((n, 1)).astype(np.float32), v, nrm], axis=1).astype(np.float32)
    raise ValueError(which)


def ctx_state(common, abi, info, kind, integral=None):
    """RtxState of a CTX_CONFIGS entry (frame 3 of the standard sequence)."""
    over = dict(environmentProb=0.0)
    if kind == "hdr":
        over = dict(common.env_state_overrides(integral))
    elif kind == "sky":
        over = dict(environmentProb=0.25)
    return common.frame_state(CTX_SIZE[0], CTX_SIZE[1], info, 3, **over)


# ---- whole trace stages (direct_stage.comp / indirect_stage.comp mains, oracle/ref_shim/ref_trace.cpp) ------------------------------
# (tag, scene maker name, (W, H), frames, environment kind, RtxState overrides)
TRACE_CONFIGS = [("cornell", "cornell_scene", (64, 40), 3, "none", dict(maxDepth=3)),
                 ("cube_hdr", "cube_scene", (48, 32), 2, "hdr", dict()),
                 ("room_sky", "small_room", (64, 48), 2, "sky", dict()),
                 ("cornell_none", "cornell_scene", (40, 24), 2, "none", dict(ReSTIRState=0, maxDepth=2)),
                 ("room_ragged", "small_room", (50, 34), 2, "none", dict(RISSampleNum=2, maxDepth=4, MIS=0)),
                 ("textured", "textured_scene", (64, 48), 2, "none", dict(maxDepth=3)),
                 ("instanced", "instanced_scene", (64, 48), 2, "none", dict(maxDepth=3)),
                 ("alpha", "alpha_scene", (64, 40), 2, "none", dict(maxDepth=3)),
                 ("spatial", "cornell_scene", (64, 40), 2, "none", dict(ReSTIRState=2, maxDepth=2)),            # eSpatial (race-free reading)
                 ("spatiotemporal", "small_room", (50, 34), 3, "none", dict(ReSTIRState=4, maxDepth=2))]      # eSpatiotemporal
# the reference's compile-time variants (eid_renderer_set_variant; abi.VARIANT_* bits): (tag, scene maker, size, frames, variant bits, RtxState overrides)
VARIANT_CONFIGS = [("bil_direct", "cornell_scene", (48, 32), 2, 1, dict(maxDepth=2)),
                   ("bil_indirect", "small_room", (48, 32), 2, 2, dict(maxDepth=3)),
                   ("sub4", "small_room", (50, 34), 3, 4, dict(maxDepth=3)),
                   ("sub4_cube_sky", "cube_scene", (40, 24), 2, 4, dict(maxDepth=2)),
                   ("all_three", "cornell_scene", (40, 24), 2, 7, dict(maxDepth=3)),
                   ("bil_nodenoise", "cornell_scene", (32, 24), 2, 3, dict(maxDepth=2, denoise=0)),
                   ("split", "cornell_scene", (48, 32), 3, 8, dict(maxDepth=2)),                 # direct_gen.comp + direct_reuse.comp
                   ("split_cube_sky", "cube_scene", (40, 24), 2, 8, dict(maxDepth=2)),
                   ("split_room_ris", "small_room", (40, 28), 2, 8 | 4, dict(maxDepth=2, ReSTIRState=1, RISSampleNum=2)),
                   ("split_debug", "cornell_scene", (32, 24), 1, 8, dict(debugging_mode=6))]
VARIANT_KEYS = ("BUF_THIS_GBUFFER", "BUF_MOTION", "BUF_THIS_DIRECT_RESV", "BUF_THIS_INDIRECT_RESV", "BUF_DIRECT", "BUF_INDIRECT", "BUF_DENOISE_DIR_A",
                "BUF_DENOISE_IND_A", "BUF_DENOISE_IND_B")
# host tables (src/scene.cpp run on an injected scene): scene makers, and the camera sequence (size, optional new look-at) after loading
SCENE_TABLE_MAKERS = ["cube_scene", "cornell_scene", "small_room", "textured_scene", "instanced_scene", "alpha_scene"]
SCENE_CAMERA_STEPS = [((640, 360), None), ((640, 360), None), ((333, 200), ((1.5, 2.5, -4.0), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0), 47.0))]
# display pass (post.frag): (tag, debugging_mode, Tonemapper overrides); rendered on DISPLAY_SCENE at DISPLAY_SIZE, DISPLAY_FRAMES frames
DISPLAY_SCENE, DISPLAY_SIZE, DISPLAY_FRAMES = "cornell_scene", (72, 44), 2
DISPLAY_CONFIGS = [("default", 0, dict()),
                   ("graded", 0, dict(brightness=1.3, contrast=0.8, saturation=0.6, vignette=0.4, avgLum=2.5)),
                   ("auto", 0, dict(autoExposure=1)),
                   ("auto_graded", 0, dict(autoExposure=1, Ywhite=0.8, key=0.3, contrast=1.2, vignette=0.2)),
                   ("direct", 1, dict()), ("direct_auto", 1, dict(autoExposure=1, key=0.7)),
                   ("indirect", 2, dict(avgLum=1.7)), ("indirect_auto", 2, dict(autoExposure=1)),
                   ("basecolor", 3, dict()), ("normal", 4, dict()),
                   ("depth", 5, dict(brightness=0.0, contrast=2.2, saturation=0.0))]       # RenderOutput::m_depthTm (render_output.hpp:56-60)
TRACE_KEYS = ("gbuffer", "motion", "direct_resv", "indirect_resv", "direct", "ind_tmp_a")


def trace_state_overrides(common, kind, integral=None, over=None):
    o = dict(environmentProb=0.0)
    if kind == "hdr":
        o = dict(common.env_state_overrides(integral))
    elif kind == "sky":
        o = dict(environmentProb=0.25)
    o.update(over or {})
    return o
