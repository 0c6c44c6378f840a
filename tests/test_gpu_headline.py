"""-m gpu: parity on the inputs bench.py times — BASELINE.json configs[1..4] at (or near) their stated sizes.

The small-scene tests of test_gpu_parity.py pin every code path; these pin the *benchmarked* inputs: the real C3 scene
(1 000 708 triangles, 1 000 emissive), the C5 stress scene (10 M triangles, 10 k lights) and C2 at 1920x1080, each over
temporal frame sequences, both K2 forms, against the CPU oracle (reference Renderer::run, renderer.cpp:154-206).
Bit-exact in strict mode; the default (fast-exp) denoiser — the variant bench.py runs — within the 1e-3 contract.
"""
import numpy as np
import pytest

import eidola_b200 as eid
from eidola_b200 import abi, scenes

import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    try:
        ok = eid.lib().eid_device_count() > 0
    except Exception:
        ok = False
    if not ok:
        pytest.fail("no CUDA device / libeidola.so: the product has no CPU fallback, GPU tests cannot run here")


@pytest.fixture(scope="module")
def c3():
    """The headline scene exactly as bench.py builds it."""
    import oracle_lib as ol
    arrays = scenes.heightfield_room(quads=707, n_light_quads=500, light_seed=566)
    osc = ol.OracleScene()
    osc.load_arrays(arrays)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    assert acc.info().triangleCount == 1000708 and psc.info().trigLightCount == 1000
    return arrays, osc, psc, acc


def _frames_vs_oracle(arrays, osc, psc, acc, size, frames, tag, forms=(1, 0), strict=True, orbit_deg=0.0, **over):
    """Oracle once per frame, every product variant (K2 wavefront / mega-kernel) against the same oracle snapshot."""
    import oracle_lib as ol
    orr = ol.OracleRenderer(osc, size)
    orr.set_env_constant(common.ENV)
    prs = []
    for form in forms:
        r = eid.Renderer()
        r.create(size, psc, acc)
        r.set_env_constant(common.ENV)
        r.set_strict_math(strict)
        r.set_wavefront(form)
        prs.append(r)
    cam = arrays.camera
    for s in (osc, psc):
        s.set_lookat(cam["eye"], cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        s.update_camera(*size)
    info = psc.info()
    worst = {}
    for f in range(frames):
        if orbit_deg:
            a = np.deg2rad(orbit_deg * f)
            e = np.array(cam["eye"], np.float64)
            eye = (e[0] * np.cos(a) - e[2] * np.sin(a), e[1], e[0] * np.sin(a) + e[2] * np.cos(a))
            for s in (osc, psc):
                s.set_lookat(eye, cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        for s in (osc, psc):
            s.update_camera(*size)
        assert osc.table(abi.TABLE_CAMERA).tobytes() == psc.table(abi.TABLE_CAMERA).tobytes()
        st = common.frame_state(size[0], size[1], info, f, **over)
        orr.run(st, f)
        want = common.snapshot(orr)
        so = orr.stats()
        for form, r in zip(forms, prs):
            r.run(st, f)
            r.sync()
            rep = common.compare_snapshots(common.snapshot(r), want, "%s form %d frame %d" % (tag, form, f))
            if strict:
                assert all(v == 0.0 for v in rep.values()), "strict math must be bit-exact, got %s" % rep
            sp = r.stats()
            assert (so.closestHitRays, so.anyHitRays, so.primaryHits) == (sp.closestHitRays, sp.anyHitRays, sp.primaryHits)
            for k, v in rep.items():
                worst[k] = max(worst.get(k, 0.0), v)
    print("%s: worst relative deviation vs oracle over %d frames: %s" % (tag, frames, worst))
    return worst


def test_c3_frames_960x540_both_k2_forms(c3):
    """C3 (BASELINE configs[2]) at a quarter of the pixels: 3 temporal frames, maxDepth 3, every buffer bit-identical."""
    arrays, osc, psc, acc = c3
    _frames_vs_oracle(arrays, osc, psc, acc, (960, 540), 3, "C3 960x540", maxDepth=3)


def test_c3_frames_1920x1080(c3):
    """C3 at the benchmarked size: two 1920x1080 frames (the second one merges temporal history)."""
    arrays, osc, psc, acc = c3
    _frames_vs_oracle(arrays, osc, psc, acc, (1920, 1080), 2, "C3 1080p", forms=(1,), maxDepth=3)


def test_c3_orbiting_camera(c3):
    """The 0.5 degree / frame orbit of SURVEY.md 8(d): reprojection really moves, reservoirs still bit-exact."""
    arrays, osc, psc, acc = c3
    _frames_vs_oracle(arrays, osc, psc, acc, (640, 360), 4, "C3 orbit", forms=(1,), orbit_deg=0.5, maxDepth=3)


def test_c3_default_fast_denoiser_within_tolerance(c3):
    """What bench.py actually times: the default denoiser (MUFU ex2 + FMA).  Integers, picks and reservoirs stay bit-exact,
    the images stay within the 1e-3 relative contract of BASELINE.json (measured ~1e-6), on the real scene at 960x540."""
    arrays, osc, psc, acc = c3
    worst = _frames_vs_oracle(arrays, osc, psc, acc, (960, 540), 3, "C3 fast-math", forms=(1,), strict=False, maxDepth=3)
    assert max(worst.values()) <= 1e-3
    assert worst["direct_resv.weight"] == 0.0 and worst["indirect_resv.weight"] == 0.0 and worst["indirect_resv.L"] == 0.0


def test_c3_spatiotemporal(c3):
    arrays, osc, psc, acc = c3
    _frames_vs_oracle(arrays, osc, psc, acc, (480, 270), 3, "C3 spatiotemporal", forms=(1,), ReSTIRState=abi.eSpatiotemporal, maxDepth=3)


def _random_rays(info, n, seed):
    rng = np.random.default_rng(seed)
    lo = np.array(info.bboxMin[:], np.float32)
    hi = np.array(info.bboxMax[:], np.float32)
    ext = hi - lo
    o = lo + rng.random((n, 3), dtype=np.float32) * ext
    t = lo + rng.random((n, 3), dtype=np.float32) * ext
    d = t - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 4:7] = o, d
    rays[:, 3] = 1e28
    return rays


def test_c3_traversal_equals_oracle_bvh_and_brute_force(c3):
    """eid_accel_trace on the 1 M-triangle BVH4: 200 000 random rays against the oracle's own (binned-SAH BVH2) intersector,
    and 3 000 of them against the brute-force loop over all 1 000 708 triangles — closest hit with tie-break and any hit."""
    import oracle_lib as ol
    arrays, osc, psc, acc = c3
    rays = _random_rays(psc.info(), 200000, 3)
    hg, ho = acc.trace(rays), osc.trace(rays)
    assert hg.tobytes() == ho.tobytes(), "closest-hit mismatch in %d rays" % int((hg != ho).sum())
    assert (hg["hitT"] < 1e27).mean() > 0.5
    brute = ol.OracleScene(use_bvh=False)
    brute.load_arrays(arrays)
    sub = rays[:3000].copy()
    hb = brute.trace(sub)
    assert hb.tobytes() == hg[:3000].tobytes(), "brute-force mismatch"
    rays[:, 3] = np.where(ho["hitT"] < 1e27, ho["hitT"] * np.float32(1.5), 5.0).astype(np.float32)
    rays[::3, 3] = (ho["hitT"][::3] * np.float32(0.5)).astype(np.float32)
    ag, ao = acc.trace(rays, any_hit=True), osc.trace(rays, any_hit=True)
    assert np.array_equal(ag["hitT"], ao["hitT"])
    assert np.array_equal(brute.trace(rays[:3000].copy(), any_hit=True)["hitT"], ag["hitT"][:3000])


def test_c2_cornell_1920x1080():
    """BASELINE configs[1] at its stated size: 32-triangle Cornell box + 2 area lights, 1920x1080, temporal DI (+ GI, denoise,
    compose), 3 frames, static camera."""
    import oracle_lib as ol
    arrays = scenes.cornell_scene()
    osc = ol.OracleScene()
    osc.load_arrays(arrays)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    _frames_vs_oracle(arrays, osc, psc, acc, (1920, 1080), 3, "C2 1080p", forms=(1,), maxDepth=3)


def test_c5_ten_million_triangles():
    """BASELINE configs[4]: the C3 generator at 2236x2236 quads (9 999 402 + 10 + 10 000 triangles, 10 000 lights): host
    tables, BVH traversal and 3 temporal frames at 640x360 (the oracle's own BVH over 10 M triangles answers its rays)."""
    import oracle_lib as ol
    arrays = scenes.heightfield_room(quads=2236, n_light_quads=5000, light_seed=567)
    osc = ol.OracleScene()
    osc.load_arrays(arrays)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    assert acc.info().triangleCount > 10_000_000 and psc.info().trigLightCount == 10000
    for t in (abi.TABLE_MATERIALS, abi.TABLE_TRIG_LIGHTS, abi.TABLE_LIGHT_INFO):
        assert osc.table(t).tobytes() == psc.table(t).tobytes()
    rays = _random_rays(psc.info(), 100000, 9)
    hg, ho = acc.trace(rays), osc.trace(rays)
    assert hg.tobytes() == ho.tobytes(), "closest-hit mismatch in %d rays" % int((hg != ho).sum())
    _frames_vs_oracle(arrays, osc, psc, acc, (640, 360), 3, "C5 640x360", maxDepth=3)


def test_c4_4k_band_of_c3(c3):
    """BASELINE configs[3] (C3 at 3840x2160): one frame is ~5 s of oracle time, so the parity check renders the 4K frame on the
    GPU and compares a 3840x272 band, which the oracle evaluates as a row band of the same 4K frame (run_trace on rows
    [1088, 1360), interior rows compared): G-buffer, motion, reservoirs and pre-denoise images of that band bit-identical."""
    import oracle_lib as ol
    arrays, osc, psc, acc = c3
    size = (3840, 2160)
    y0, y1 = 1088, 1360
    m = 16      # rows next to the band edge may reproject into rows the oracle never rendered: compare the interior
    orr = ol.OracleRenderer(osc, size)
    orr.set_env_constant(common.ENV)
    r = eid.Renderer()
    r.create(size, psc, acc)
    r.set_env_constant(common.ENV)
    cam = arrays.camera
    for s in (osc, psc):
        s.set_lookat(cam["eye"], cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        s.update_camera(*size)
    info = psc.info()
    for f in range(2):
        for s in (osc, psc):
            s.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, maxDepth=3)
        orr.run_trace(st, f, y0, y1)
        r.run_trace(st, f)
        r.sync()
        for name, which in (("gbuffer", abi.BUF_THIS_GBUFFER), ("motion", abi.BUF_MOTION), ("direct_resv", abi.BUF_THIS_DIRECT_RESV),
                            ("direct", abi.BUF_DIRECT)):
            g, o = r.read(which), orr.read(which)
            rows_g = np.ascontiguousarray(g).view(np.uint8).reshape(size[1], -1)[y0 + m:y1 - m]
            rows_o = np.ascontiguousarray(o).view(np.uint8).reshape(size[1], -1)[y0 + m:y1 - m]
            assert rows_g.tobytes() == rows_o.tobytes(), "C4 band: %s differs in frame %d" % (name, f)
        for name, which in (("indirect_resv", abi.BUF_THIS_INDIRECT_RESV),):
            g, o = r.read(which), orr.read(which)
            rows_g = np.ascontiguousarray(g).view(np.uint8).reshape(size[1] // 2, -1)[(y0 + m) // 2:(y1 - m) // 2]
            rows_o = np.ascontiguousarray(o).view(np.uint8).reshape(size[1] // 2, -1)[(y0 + m) // 2:(y1 - m) // 2]
            assert rows_g.tobytes() == rows_o.tobytes(), "C4 band: %s differs in frame %d" % (name, f)
