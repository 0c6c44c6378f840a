"""-m gpu: parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): G-buffer words, motion indices, reservoir `num` and the picked samples
bit-exact; weights and radiance within 1e-3 relative (absolute floor 1e-6).  With the shared numerical
contract the deviations are expected to be exactly 0; the tests print what they measured.
"""
import os

import numpy as np
import pytest

import eidola_b200 as eid
from eidola_b200 import abi, scenes

import common

pytestmark = pytest.mark.gpu


def _have_gpu():
    try:
        return eid.lib().eid_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not _have_gpu():
        pytest.fail("no CUDA device / libeidola.so: the product has no CPU fallback, GPU tests cannot run here")


def _random_rays(info, n, seed):
    rng = np.random.default_rng(seed)
    lo = np.array(info.bboxMin[:], np.float32)
    hi = np.array(info.bboxMax[:], np.float32)
    ext = hi - lo
    o = lo - 0.1 * ext + rng.random((n, 3), dtype=np.float32) * 1.2 * ext
    t = lo + rng.random((n, 3), dtype=np.float32) * ext
    d = t - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 4:7] = o, d
    rays[:, 3] = 1e28
    # axis-parallel and degenerate directions too
    rays[0, 4:7] = (1, 0, 0)
    rays[1, 4:7] = (0, -1, 0)
    rays[2, 4:7] = (0, 0, 0)
    return rays


@pytest.mark.parametrize("maker", [scenes.cube_scene, scenes.cornell_scene, scenes.small_room])
def test_tables_match_oracle(maker):
    import oracle_lib as ol
    arrays = maker()
    o = ol.OracleScene()
    o.load_arrays(arrays)
    p = eid.Scene(0)
    p.load_arrays(arrays)
    for t in (abi.TABLE_MATERIALS, abi.TABLE_PUNC_LIGHTS, abi.TABLE_TRIG_LIGHTS, abi.TABLE_LIGHT_INFO):
        assert o.table(t).tobytes() == p.table(t).tobytes()
    inst = p.table(abi.TABLE_INSTANCE_DATA)
    assert np.array_equal(inst["materialIndex"], o.table(abi.TABLE_INSTANCE_DATA)["materialIndex"])
    assert (inst["vertexAddress"] != 0).all() and (inst["indexAddress"] != 0).all()   # real device addresses
    for pm in range(p.info().primMeshCount):
        assert o.table(abi.TABLE_VERTICES, pm).tobytes() == p.table(abi.TABLE_VERTICES, pm).tobytes()
        assert o.table(abi.TABLE_INDICES, pm).tobytes() == p.table(abi.TABLE_INDICES, pm).tobytes()


@pytest.mark.parametrize("maker,n", [(scenes.cube_scene, 20000), (scenes.cornell_scene, 50000), (scenes.small_room, 100000)])
def test_traversal_equals_brute_force(maker, n):
    """CUDA BVH traversal == the oracle's brute-force loop over every triangle (closest hit with tie-break, any hit)."""
    import oracle_lib as ol
    arrays = maker()
    o = ol.OracleScene(use_bvh=False)
    o.load_arrays(arrays)
    p = eid.Scene(0)
    p.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(p)
    rays = _random_rays(p.info(), n, 11)
    hg, ho = acc.trace(rays), o.trace(rays)
    assert hg.tobytes() == ho.tobytes(), "closest-hit mismatch in %d rays" % int((hg != ho).sum())
    assert (hg["hitT"] < 1e27).mean() > 0.3
    rays[:, 3] = np.where(ho["hitT"] < 1e27, ho["hitT"] * np.float32(1.5), 5.0).astype(np.float32)
    rays[::3, 3] = (ho["hitT"][::3] * np.float32(0.5)).astype(np.float32)   # tmax short of the hit: must be unoccluded
    ag, ao = acc.trace(rays, any_hit=True), o.trace(rays, any_hit=True)
    assert np.array_equal(ag["hitT"], ao["hitT"])


def test_oracle_bvh_equals_brute_force():
    """Sanity of the checker itself: the oracle's BVH path and its brute-force path agree."""
    import oracle_lib as ol
    arrays = scenes.small_room()
    a, b = ol.OracleScene(use_bvh=True), ol.OracleScene(use_bvh=False)
    a.load_arrays(arrays)
    b.load_arrays(arrays)
    rays = _random_rays(a.info(), 50000, 5)
    assert a.trace(rays).tobytes() == b.trace(rays).tobytes()


def _run_frames(arrays, size, frames, tag, orbit=False, strict=True, env_img=None, wavefront=True, sun_sky=None, **state_over):
    osc, orr, psc, acc, prr = common.make_pair(arrays, size, strict=strict, env_img=env_img)
    prr.set_wavefront(wavefront)
    if sun_sky is not None:
        orr.set_sun_and_sky(sun_sky)
        prr.set_sun_and_sky(sun_sky)
    if env_img is not None:
        state_over = dict(common.env_state_overrides(prr._env.get_integral()), **state_over)
    for s in (osc, psc):
        s.update_camera(*size)     # contract: one updateCamera before frame 0 so last* matrices are valid
    info = psc.info()
    cam = arrays.camera
    worst = {}
    for f in range(frames):
        if orbit:
            a = np.deg2rad(0.5 * f)
            e = np.array(cam["eye"], np.float64)
            eye = (e[0] * np.cos(a) - e[2] * np.sin(a), e[1], e[0] * np.sin(a) + e[2] * np.cos(a))
            for s in (osc, psc):
                s.set_lookat(eye, cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        for s in (osc, psc):
            s.update_camera(*size)
        assert osc.table(abi.TABLE_CAMERA).tobytes() == psc.table(abi.TABLE_CAMERA).tobytes()
        st = common.frame_state(size[0], size[1], info, f, **state_over)
        orr.run(st, f)
        prr.run(st, f)
        prr.sync()
        rep = common.compare_snapshots(common.snapshot(prr), common.snapshot(orr), "%s frame %d" % (tag, f))
        for k, v in rep.items():
            worst[k] = max(worst.get(k, 0.0), v)
        if strict:
            assert all(v == 0.0 for v in rep.values()), "strict math must be bit-exact, got %s" % rep
        so, sp = orr.stats(), prr.stats()
        assert (so.closestHitRays, so.anyHitRays, so.primaryHits) == (sp.closestHitRays, sp.anyHitRays, sp.primaryHits)
    print("%s: worst relative deviation vs oracle over %d frames: %s" % (tag, frames, worst))
    return worst


@pytest.mark.parametrize("restir", [abi.eSpatial, abi.eSpatiotemporal])
def test_spatial_reuse_matches_oracle(restir):
    """eSpatial / eSpatiotemporal (direct_stage.comp:224-255) in its race-free reading (every tempDirectResv write before any read,
    DESIGN.md §3): k_direct_stage<SPATIAL> + k_direct_spatial against the oracle — which the reference's own shader text, dispatched
    twice, reproduces bit for bit (tests/golden/ref_trace.npz, configurations `spatial` / `spatiotemporal`).  Covers sky pixels and
    emitters (they never write their tempDirectResv entry), a frame smaller than the allocation, an orbiting camera, both K2 forms."""
    for arrays, size, frames, tag, kw in ((scenes.small_room(), (256, 144), 4, "room", dict(orbit=True)),
                                           (scenes.cube_scene(), (96, 64), 3, "cube", dict(wavefront=False)),
                                           (scenes.cornell_scene(), (200, 120), 3, "cornell", dict(RISSampleNum=2, maxDepth=2))):
        worst = _run_frames(arrays, size, frames, "%s-restir%d" % (tag, restir), ReSTIRState=restir, **kw)
        assert max(worst.values()) == 0.0
    # tempDirectResv itself, and a frame smaller than the allocation (pitch of the reservoir buffers = size.x, of the images = allocation)
    arrays = scenes.small_room()
    osc, orr, psc, acc, prr = common.make_pair(arrays, (160, 96))
    for s in (osc, psc):
        s.update_camera(120, 80)
    info = psc.info()
    for f in range(3):
        for s in (osc, psc):
            s.update_camera(120, 80)
        st = common.frame_state(120, 80, info, f, ReSTIRState=restir)
        orr.run(st, f); prr.run(st, f); prr.sync()
        rep = common.compare_snapshots(common.snapshot(prr), common.snapshot(orr), "sub-allocation frame %d" % f)
        assert all(v == 0.0 for v in rep.values()), rep
        assert prr.read(abi.BUF_TEMP_DIRECT_RESV).tobytes() == orr.read(abi.BUF_TEMP_DIRECT_RESV).tobytes()
    # the same renderer back on temporal-only reuse: the spatial scratch is simply not touched
    st = common.frame_state(120, 80, info, 3)
    orr.run(st, 3); prr.run(st, 3); prr.sync()
    assert all(v == 0.0 for v in common.compare_snapshots(common.snapshot(prr), common.snapshot(orr), "back to temporal").values())


def test_fast_math_denoiser_within_tolerance():
    """Default (MUFU ex2) denoiser: images within the 1e-3 contract (measured ~1e-6); ints and reservoirs stay bit-exact."""
    worst = _run_frames(scenes.small_room(), (256, 144), 3, "room-fastmath", strict=False)
    assert max(worst.values()) <= 1e-3
    assert worst["direct_resv.weight"] == 0.0 and worst["indirect_resv.weight"] == 0.0 and worst["indirect_resv.L"] == 0.0


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("size", [(256, 144), (130, 70), (64, 2), (700, 300)])
def test_denoiser_kernel_forms_agree(strict, size):
    """The A-Trous passes exist as a shared-memory tile kernel (tiles by TMA = default, or by cp.async; 2 or 4 lattice rows per
    thread) and as the round-1 kernel (taps through L1; 4, 2 or 1 pixels per thread).  Every form gives each pixel its taps in the
    reference's order, so with strict math ALL forms are bit-identical; with the default numerics the forms of one kernel are
    bit-identical among themselves and the two kernels agree to ~1e-6 (different but equivalent fast arithmetic)."""
    arrays = scenes.small_room()
    snaps = []
    forms = [(1, 4), (1, 2), (2, 4), (2, 2), (0, 4), (0, 2), (0, 1)]
    for mode, rows in forms:
        osc, orr, psc, acc, prr = common.make_pair(arrays, size, strict=strict)
        if mode:
            prr.set_denoise_tiles(mode, rows)
        else:
            prr.set_denoise_tiles(0)
            prr.set_denoise_rows(rows)
        psc.update_camera(*size)
        for f in range(3):
            psc.update_camera(*size)
            prr.run(common.frame_state(size[0], size[1], psc.info(), f), f)
        prr.sync()
        snaps.append(common.snapshot(prr))
    for name in ("direct", "indirect", "ind_tmp_a", "ind_tmp_b"):
        for k in (1, 2, 3):
            assert snaps[k][name].tobytes() == snaps[0][name].tobytes(), "tile kernel forms differ in %s (%s)" % (name, forms[k])
        for k in (5, 6):
            assert snaps[k][name].tobytes() == snaps[4][name].tobytes(), "legacy kernel forms differ in %s (%s)" % (name, forms[k])
        if strict:
            assert snaps[4][name].tobytes() == snaps[0][name].tobytes(), "tile kernel != legacy kernel in strict mode (%s)" % name
        else:
            assert common.rel_err(snaps[4][name], snaps[0][name]) < 1e-4, name
    with pytest.raises(eid.EidolaError):
        prr.set_denoise_rows(3)
    with pytest.raises(eid.EidolaError):
        prr.set_denoise_tiles(1, 3)


@pytest.mark.parametrize("maker,size", [(scenes.cube_scene, (192, 192)), (scenes.cornell_scene, (224, 128))])
def test_hdr_environment_default_state(maker, size):
    """Scope row (f.2): HDR environment importance sampling with the reference's DEFAULT RtxState (environmentProb 0.25):
    EnvSample / Environment_sample / EnvRadiance / EnvEval + alias map, open scenes where the sky is visible and lights the GI."""
    worst = _run_frames(maker(), size, 4, "hdr-env " + maker.__name__, env_img=scenes.synthetic_sky())
    assert max(worst.values()) == 0.0
    _run_frames(maker(), size, 2, "hdr-env no sun, prob 0.6", env_img=scenes.synthetic_sky(sun=False), environmentProb=0.6, hdrMultiplier=2.0)


@pytest.mark.parametrize("over", [dict(), dict(ReSTIRState=abi.eNone, denoise=0), dict(maxDepth=2, MIS=0)])
@pytest.mark.parametrize("wavefront", [True, False])
def test_sun_and_sky_environment(over, wavefront):
    """SunAndSky.in_use = 1 (shaders/sun_and_sky.glsl): EnvRadiance on primary misses, EnvSample (two draws inside the sun's glow
    disc, pdf 0.5) for 25 % of the light candidates, EnvEval on bounce misses; defaults of sample_example.hpp:186-203 and a low sun."""
    for ss in (abi.default_sun_and_sky(in_use=1),
               abi.default_sun_and_sky(in_use=1, sun_direction=abi.Vec3(0.6, 0.08, 0.35), haze=3.0, redblueshift=0.1, saturation=1.3, horizon_height=0.5)):
        worst = _run_frames(scenes.cube_scene(), (160, 128), 3, "sun&sky %s" % over, sun_sky=ss, wavefront=wavefront,
                            environmentProb=0.25, fireflyClampThreshold=50.0, **over)
        assert max(worst.values()) == 0.0


@pytest.mark.parametrize("mode", [abi.eNoDebug, abi.eDirectStage, abi.eIndirectStage, abi.eBaseColor, abi.eDepth])
def test_display_pass_post_frag(mode):
    """RenderOutput::run -> post.frag (tonemap, dither, contrast / brightness / saturation / vignette) on the frame rendered last:
    float output bit-identical to the oracle's restatement, RGBA8 = round-half-away packing of it."""
    size = (200, 120)
    osc, orr, psc, acc, prr = common.make_pair(scenes.cornell_scene(), size)
    for s_ in (osc, psc):
        s_.update_camera(*size)
    info = psc.info()
    for f in range(2):
        for s_ in (osc, psc):
            s_.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, maxDepth=3, debugging_mode=mode)
        orr.run(st, f); prr.run(st, f)
    for tm in (abi.default_tonemapper(), abi.default_tonemapper(brightness=1.3, contrast=0.8, saturation=0.6, vignette=0.4, avgLum=2.5),
               abi.default_tonemapper(autoExposure=1), abi.default_tonemapper(autoExposure=1, Ywhite=0.9, key=0.25, vignette=0.3)):   # auto exposure: 1x1 level of the blit chain
        if mode == abi.eDepth:
            tm = abi.default_tonemapper(brightness=0.0, contrast=2.2, saturation=0.0)     # RenderOutput::m_depthTm (render_output.hpp:56-60)
        want = orr.run_output(tm, st)
        prr.run_output(tm); prr.sync()
        got = prr.read(abi.BUF_DISPLAY_F32).reshape(size[1], size[0], 4)
        assert np.isfinite(want[..., :3]).all() or mode == abi.eDepth or tm.autoExposure      # toneExposure of a black pixel is 0/0, there as here
        assert common.same_bits_or_both_nan(got, want), "display pass differs from the oracle (max abs %g)" % np.nanmax(np.abs(got - want))
        got8 = prr.read(abi.BUF_DISPLAY_RGBA8).reshape(size[1], size[0], 4)
        c = np.nan_to_num(np.clip(want.astype(np.float32), 0.0, 1.0), nan=0.0)
        want8 = np.floor(c * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
        assert np.array_equal(got8, want8)
    with pytest.raises(eid.EidolaError):
        prr.run_output(abi.default_tonemapper(autoExposure=3))      # toneLocalExposure: never selected by the reference's GUI, outside the contract


def test_cuda_display_pass_matches_reference_post_frag():
    """k_mip_blit + k_post against the committed output of the reference's OWN post.frag (main() included, compiled as C++ by
    oracle/ref_shim/ref_display.cpp; tests/golden/ref_display.npz): every view and tonemapper of DISPLAY_CONFIGS, auto exposure included."""
    import ref_fn_inputs as fi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_display.npz"))
    size = fi.DISPLAY_SIZE
    rendered = {}
    for tag, mode, over in fi.DISPLAY_CONFIGS:
        if mode not in rendered:
            psc = eid.Scene(0); psc.load_arrays(getattr(scenes, fi.DISPLAY_SCENE)())
            acc = eid.AccelStructure(); acc.create(psc)
            prr = eid.Renderer(); prr.create(size, psc, acc); prr.set_env_constant(common.ENV); prr.set_strict_math(True)
            psc.update_camera(*size)
            info = psc.info()
            for f in range(fi.DISPLAY_FRAMES):
                psc.update_camera(*size)
                prr.run(common.frame_state(size[0], size[1], info, f, maxDepth=3, debugging_mode=mode), f)
            rendered[mode] = (psc, acc, prr)
        prr = rendered[mode][2]
        prr.run_output(abi.default_tonemapper(**over)); prr.sync()
        got = prr.read(abi.BUF_DISPLAY_F32).reshape(size[1], size[0], 4)
        assert common.same_bits_or_both_nan(got, z["%s_out" % tag]), tag


def test_device_functions_match_reference_glsl_vectors():
    """The device-side shader functions and sun_and_sky() against the committed outputs of the reference's OWN GLSL text compiled as C++
    (tests/golden/ref_vectors.npz, made by oracle/ref_shim from /root/reference): bit for bit."""
    import ctypes as C
    import ref_fn_inputs as fi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
    L = eid.lib()
    for w, (ni, no) in enumerate(fi.ARITY):
        if w in (9, 10):
            continue                      # reservoir operations are inline in the stage kernels (frame-level parity covers them)
        x = np.ascontiguousarray(fi.inputs(w, n=1500))
        got = np.zeros((x.shape[0], no), np.float32)
        assert L.eid_fn_tap(0, w, x.ctypes.data, x.shape[0], got.ctypes.data) == 0
        want = z["fn_%d_out" % w]
        bad = np.nonzero((got.view(np.uint32) != want.view(np.uint32)).any(axis=1))[0]
        assert bad.size == 0, "%s: %d of %d items differ from the reference GLSL, first: in %s got %s want %s" % (
            fi.NAMES[w], bad.size, len(got), x[bad[0]], got[bad[0]], want[bad[0]])
    dirs = np.ascontiguousarray(z["sky_dirs"])
    for k, kw in enumerate(fi.SKY_PARAMS):
        ss = fi.sun_sky(abi, kw)
        got = np.zeros_like(dirs)
        assert L.eid_sun_and_sky_eval(0, C.byref(ss), dirs.ctypes.data, len(dirs), got.ctypes.data) == 0
        assert got.view(np.uint32).tobytes() == z["sky_%d_out" % k].view(np.uint32).tobytes(), "sun_and_sky, parameter set %d" % k


def test_cuda_trace_stages_match_reference_shader_mains():
    """The CUDA direct / indirect stages (both forms of the indirect stage) against the committed output of the reference's OWN
    direct_stage.comp / indirect_stage.comp (main() included, compiled as C++ by oracle/ref_shim/ref_trace.cpp): G-buffer, motion vectors,
    reservoirs and pre-denoise images after the last frame of every configuration, bit for bit."""
    import ref_fn_inputs as fi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_trace.npz"))
    for c in fi.TRACE_CONFIGS:
        tag, maker_name, size, frames, kind, over = c
        for wavefront in (1, 0):
            psc = eid.Scene(0); psc.load_arrays(getattr(scenes, maker_name)())
            acc = eid.AccelStructure(); acc.create(psc)
            prr = eid.Renderer(); prr.create(size, psc, acc); prr.set_env_constant((0.0, 0.0, 0.0)); prr.set_strict_math(True); prr.set_wavefront(wavefront)
            integral = None
            if kind == "hdr":
                penv = eid.HdrSampling(0); penv.set_pixels(fi.ctx_env_image(scenes)); prr.set_env(penv); integral = penv.get_integral()
            elif kind == "sky":
                prr.set_sun_and_sky(fi.sun_sky(abi, fi.CTX_SKY))
            o = fi.trace_state_overrides(common, kind, integral, over)
            psc.update_camera(*size)
            info = psc.info()
            for f in range(frames):
                psc.update_camera(*size)
                st = common.frame_state(size[0], size[1], info, f, **o)
                if f < frames - 1:
                    prr.run(st, f)
                else:
                    prr.run_trace(st, f)          # the last frame stops before the denoisers: the pre-denoise images are compared
            prr.sync()
            got = {"gbuffer": prr.read(abi.BUF_THIS_GBUFFER), "motion": prr.read(abi.BUF_MOTION), "direct_resv": prr.read(abi.BUF_THIS_DIRECT_RESV),
                   "indirect_resv": prr.read(abi.BUF_THIS_INDIRECT_RESV), "direct": prr.read(abi.BUF_DIRECT), "ind_tmp_a": prr.read(abi.BUF_DENOISE_IND_A)}
            for k in fi.TRACE_KEYS:
                assert np.ascontiguousarray(got[k]).view(np.uint8).reshape(-1).tobytes() == z["%s_%s" % (tag, k)].tobytes(), (tag, k, "wavefront" if wavefront else "mega")


def test_cuda_post_stages_match_reference_shader_mains():
    """The CUDA denoisers + compose (strict math) against the committed output of the reference's OWN denoise_direct.comp x4,
    denoise_indirect.comp x5, compose.comp (main() included, compiled as C++ by oracle/ref_shim/ref_post.cpp): bit for bit."""
    import make_golden_cfg as cfg
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_post.npz"))
    for name in ("c2_cornell", "room"):
        maker, size, frames, over = cfg.CONFIGS[name][:4]
        psc = eid.Scene(0); psc.load_arrays(maker())
        acc = eid.AccelStructure(); acc.create(psc)
        prr = eid.Renderer(); prr.create(size, psc, acc); prr.set_env_constant(common.ENV); prr.set_strict_math(True)
        psc.update_camera(*size)
        info = psc.info()
        for f in range(frames):
            psc.update_camera(*size)
            prr.run(common.frame_state(size[0], size[1], info, f, **over), f)
        prr.sync()
        for k in ("BUF_DIRECT", "BUF_INDIRECT", "BUF_DENOISE_IND_A", "BUF_DENOISE_IND_B"):
            assert prr.read(getattr(abi, k)).tobytes() == z["%s_%s" % (name, k)].tobytes(), (name, k)


def test_device_light_sampling_matches_reference_glsl_vectors():
    """The device's light sampling (triangle / punctual lights, HDR alias map, sun & sky), EnvEval, EnvRadiance, raySpawn, clampRadiance and
    Sample, evaluated with the renderer's own scene tables, against the committed outputs of pathtrace.glsl / env_sampling.glsl compiled
    as C++ (tests/golden/ref_vectors.npz): bit for bit, including the RNG state after the draws."""
    import ctypes as C
    import ref_fn_inputs as fi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
    for tag, maker_name, kind in fi.CTX_CONFIGS:
        psc = eid.Scene(0); psc.load_arrays(getattr(scenes, maker_name)())
        acc = eid.AccelStructure(); acc.create(psc)
        rr = eid.Renderer(); rr.create(fi.CTX_SIZE, psc, acc); rr.set_env_constant(common.ENV)
        psc.update_camera(*fi.CTX_SIZE); psc.update_camera(*fi.CTX_SIZE)
        integral = None
        if kind == "hdr":
            penv = eid.HdrSampling(0); penv.set_pixels(fi.ctx_env_image(scenes)); rr.set_env(penv); integral = penv.get_integral()
        elif kind == "sky":
            rr.set_sun_and_sky(fi.sun_sky(abi, fi.CTX_SKY))
        st = fi.ctx_state(common, abi, psc.info(), kind, integral)
        nmat = len(psc.table(abi.TABLE_MATERIALS))
        for w, (ni, no) in enumerate(fi.CTX_ARITY):
            key = "ctx_%s_%d_out" % (tag, w)
            if w == 1 or key not in z.files:
                continue
            x = np.ascontiguousarray(fi.ctx_inputs(w, nmat))
            got = np.zeros((x.shape[0], no), np.float32)
            assert eid.lib().eid_renderer_fn_tap(rr._h, C.byref(st), w, x.ctypes.data, x.shape[0], got.ctypes.data) == 0
            want = z[key]
            bad = np.nonzero((got.view(np.uint32) != want.view(np.uint32)).any(axis=1))[0]
            assert bad.size == 0, "%s / %s: %d of %d items differ from the reference GLSL, first: in %s got %s want %s" % (
                tag, fi.CTX_NAMES[w], bad.size, len(got), x[bad[0]], got[bad[0]], want[bad[0]])


def test_sun_and_sky_function_matches_oracle():
    """sun_and_sky(ss, dir) on the device == the oracle's restatement, bit for bit, over random directions and parameter sets."""
    import ctypes as C
    import oracle_lib as ol
    rng = np.random.default_rng(11)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    d = np.ascontiguousarray(d / np.linalg.norm(d, axis=1, keepdims=True))
    d[:2000] = (np.array([0.0, 0.78, 0.62], np.float32) / np.float32(0.99639) + 0.05 * d[:2000]).astype(np.float32)   # around the sun
    d = np.ascontiguousarray(d / np.linalg.norm(d, axis=1, keepdims=True))
    for ss in (abi.default_sun_and_sky(in_use=1), abi.default_sun_and_sky(in_use=1, haze=5.0, saturation=1.4, redblueshift=-0.2, horizon_height=-0.7, horizon_blur=0.0),
               abi.default_sun_and_sky(in_use=1, sun_direction=abi.Vec3(0.7, -0.1, 0.2), physically_scaled_sun=0, y_is_up=0)):
        want = np.zeros_like(d); got = np.zeros_like(d)
        ol.lib().orc_sun_and_sky(C.byref(ss), d.ctypes.data, len(d), want.ctypes.data)
        assert eid.lib().eid_sun_and_sky_eval(0, C.byref(ss), d.ctypes.data, len(d), got.ctypes.data) == 0
        bad = np.nonzero((got.view(np.uint32) != want.view(np.uint32)).any(axis=1))[0]
        assert bad.size == 0, "%d of %d directions differ, first: dir %s got %s want %s" % (bad.size, len(d), d[bad[0]], got[bad[0]], want[bad[0]])


def test_sun_and_sky_off_needs_an_environment():
    psc = eid.Scene(0); psc.load_arrays(scenes.cube_scene())
    acc = eid.AccelStructure(); acc.create(psc)
    rr = eid.Renderer(); rr.create((64, 64), psc, acc)
    psc.update_camera(64, 64)
    st = common.frame_state(64, 64, psc.info(), 0, environmentProb=0.25)
    with pytest.raises(eid.EidolaError):
        rr.run(st, 0)
    rr.set_sun_and_sky(abi.default_sun_and_sky(in_use=1))
    rr.run(st, 0); rr.sync()
    assert np.isfinite(rr.read(abi.BUF_DIRECT)).all()


def test_textured_materials_normal_maps_and_textured_emitters():
    """Scope row (f.1), texture half: every textureLod tap of GetMaterials / LightEval / SampleTriangleLight (base colour with
    sRGB decode, metallic-roughness, normal map + tangent frame, emissive map, transmission map), NEAREST/LINEAR filters and
    REPEAT / MIRRORED_REPEAT / CLAMP_TO_EDGE wrap modes, default-white fallbacks."""
    worst = _run_frames(scenes.textured_scene(), (256, 160), 4, "textured")
    assert max(worst.values()) == 0.0
    _run_frames(scenes.textured_scene(), (160, 96), 2, "textured-env", env_img=scenes.synthetic_sky(), maxDepth=3)


def test_stochastic_alpha_mask_and_blend():
    """Scope row (f.1), alpha half: HitTest (traceray_rq.glsl:32-102) on MASK / BLEND instances for primary, shadow and bounce
    rays, with the candidate order pinned to front-to-back (t, instanceID, primitiveID); one RNG draw per tested candidate."""
    worst = _run_frames(scenes.alpha_scene(), (256, 160), 4, "alpha")
    assert max(worst.values()) == 0.0
    _run_frames(scenes.alpha_scene(), (128, 96), 2, "alpha-eNone", ReSTIRState=abi.eNone, maxDepth=2)


def test_instanced_and_mirrored_nodes():
    """One prim mesh instanced by three nodes (rotation + non-uniform scale, a mirroring transform), an emissive triangle on a transformed
    node: GetState's object-to-world handling, facing decided in object space, world-space TrigLight vertices."""
    worst = _run_frames(scenes.instanced_scene(), (192, 128), 3, "instanced", maxDepth=3)
    assert max(worst.values()) == 0.0
    _run_frames(scenes.instanced_scene(), (96, 64), 2, "instanced-mega", wavefront=False, maxDepth=3)


def test_c1_cube_direct_only():
    """BASELINE config 0: single cube, 256x256, RIS M=1, no denoise, direct only."""
    _run_frames(scenes.cube_scene(), (256, 256), 1, "C1", ReSTIRState=abi.eRIS, RISSampleNum=1, denoise=0)


def test_c2_cornell_temporal_di_gi():
    """BASELINE config 1 scene at a size the oracle renders in seconds: temporal DI + GI + denoise, 4 frames."""
    _run_frames(scenes.cornell_scene(), (320, 180), 4, "C2", maxDepth=3)


def test_small_room_full_pipeline_static_and_orbit():
    """Miniature of the headline C3 scene (noise height-field room, emissive quads): full pipeline, depth 4."""
    _run_frames(scenes.small_room(), (256, 144), 3, "room-static")
    _run_frames(scenes.small_room(), (256, 144), 3, "room-orbit", orbit=True)


@pytest.mark.parametrize("over", [dict(ReSTIRState=abi.eNone), dict(ReSTIRState=abi.eRIS, RISSampleNum=8), dict(MIS=0),
                                  dict(modulate=0), dict(denoise=0), dict(maxDepth=1), dict(debugging_mode=abi.eNormal),
                                  dict(reservoirClamp=2), dict(fireflyClampThreshold=0.5)])
def test_state_variants(over):
    _run_frames(scenes.cornell_scene(), (160, 96), 3, "variant %s" % over, **over)


@pytest.mark.parametrize("over", [dict(), dict(maxDepth=1), dict(maxDepth=6), dict(MIS=0), dict(ReSTIRState=abi.eRIS)])
def test_indirect_stage_megakernel_form(over):
    """indirect_stage has two forms (ray queues + dynamic-fetch traversal, default; one thread per pixel): the default form is
    what every other test runs, this one pins the one-thread-per-pixel kernel to the oracle as well."""
    _run_frames(scenes.small_room(), (192, 112), 3, "megakernel %s" % over, wavefront=False, **over)
    _run_frames(scenes.small_room(), (192, 112), 2, "wavefront %s" % over, wavefront=True, **over)


def test_indirect_stage_forms_bit_identical_large():
    """Both forms of indirect_stage on a frame too large for the oracle: every buffer bit-identical over 4 frames (moving camera),
    for several grids of the persistent traversal kernel."""
    arrays = scenes.small_room()
    size = (960, 544)
    psc = eid.Scene(0); psc.load_arrays(arrays)
    acc = eid.AccelStructure(); acc.create(psc)
    rrs = []
    for on, blocks in ((0, 0), (1, 0), (2, 0), (1, 7), (1, 2000)):
        rr = eid.Renderer(); rr.create(size, psc, acc); rr.set_env_constant(common.ENV); rr.set_wavefront(on, blocks)
        rrs.append(rr)
    info = psc.info()
    cam = arrays.camera
    psc.update_camera(*size)
    for f in range(4):
        a = np.deg2rad(0.7 * f)
        e = np.array(cam["eye"], np.float64)
        psc.set_lookat((e[0] * np.cos(a) - e[2] * np.sin(a), e[1], e[0] * np.sin(a) + e[2] * np.cos(a)), cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, maxDepth=4)
        snaps = []
        for rr in rrs:
            rr.run(st, f); rr.sync()
            snaps.append(common.snapshot(rr))
            s = rr.stats()
            snaps[-1]["rays"] = np.array([s.closestHitRays, s.anyHitRays, s.primaryHits])
        for other in snaps[1:]:
            for k in snaps[0]:
                assert np.asarray(snaps[0][k]).tobytes() == np.asarray(other[k]).tobytes(), "frame %d: %s differs between the two forms" % (f, k)


@pytest.mark.parametrize("size", [(8, 8), (17, 9), (130, 70), (64, 2)])
def test_ragged_sizes(size):
    """Odd / tiny sizes: partial 8x8 tiles, W/2 truncation, bottom tile row that relies on the early-return."""
    _run_frames(scenes.cornell_scene(), size, 2, "size %dx%d" % size, maxDepth=2)


def _tiny_scene(n_tris, with_light=True, emissive=False):
    _Builder, IDENTITY = scenes._Builder, scenes.IDENTITY
    b = _Builder()
    mat = b.add_material(base=(0.7, 0.6, 0.5, 1.0), metallic=0.2, roughness=0.6, **(dict(emissive=(4.0, 3.0, 2.0)) if emissive else {}))
    tris = [[(-1.0, -0.5 + 0.3 * k, 0.2 * k), (1.0, -0.5 + 0.3 * k, 0.2 * k), (0.0, 0.8 + 0.1 * k, 0.2 * k + 0.1)] for k in range(n_tris)]
    if n_tris:
        b.add_tris(np.array(tris)[:, ::-1], mat)      # facing the camera at -z
    if with_light:
        m = list(IDENTITY); m[12], m[13], m[14] = 0.5, 2.0, -3.0
        b.lights.append(dict(worldMatrix=m, type=1, color=(1.0, 0.9, 0.8), intensity=20.0))
    cam = dict(eye=(0.2, 0.4, -4.0), center=(0.0, 0.2, 0.0), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(50.0)))
    return b.build(cam, "tiny%d" % n_tris)


@pytest.mark.parametrize("wavefront", [True, False])
@pytest.mark.parametrize("n_tris,with_light,emissive", [(1, True, False), (2, False, False), (3, False, True), (5, True, True)])
def test_degenerate_scenes(n_tris, with_light, emissive, wavefront):
    """One-leaf BVHs (the root reference is a leaf), scenes without any light (every light sample is InvalidPdf), emissive-only scenes:
    the queue traversal and both K2 forms against the oracle."""
    _run_frames(_tiny_scene(n_tris, with_light, emissive), (96, 64), 3, "tiny %d %s %s" % (n_tris, with_light, emissive), wavefront=wavefront, maxDepth=3)


@pytest.mark.parametrize("over", [dict(maxDepth=0), dict(RISSampleNum=0), dict(maxDepth=25), dict(maxDepth=26)])
def test_extreme_state_values(over):
    """maxDepth 0 (no bounce at all), no light candidates, the deepest wavefront depth (25) and the first one that falls back to the
    one-thread-per-pixel kernel (26)."""
    _run_frames(scenes.cornell_scene(), (64, 48), 2, "extreme %s" % over, **over)


def test_state_size_smaller_than_allocation():
    """De-scaling (sample_example.cpp:396-401): RtxState.size below the allocated size; reservoirs are pitched by size.x."""
    arrays = scenes.cornell_scene()
    osc, orr, psc, acc, prr = common.make_pair(arrays, (200, 120))
    info = psc.info()
    for f in range(3):
        for s in (osc, psc):
            s.update_camera(100, 60)
        st = common.frame_state(100, 60, info, f)
        orr.run(st, f)
        prr.run(st, f)
        prr.sync()
        rep = common.compare_snapshots(common.snapshot(prr), common.snapshot(orr), "descaled frame %d" % f)
        assert all(v == 0.0 for v in rep.values()), "strict math must be bit-exact, got %s" % rep


@pytest.mark.parametrize("restir,exchange_history,orbit", [(abi.eTemporal, True, False), (abi.eSpatiotemporal, True, False), (abi.eSpatiotemporal, False, False),
                                                          (abi.eTemporal, True, True), (abi.eSpatiotemporal, True, True)])
def test_band_sharded_trace_equals_full_frame(restir, exchange_history, orbit):
    """Multi-GPU decomposition on one device: two renderers trace disjoint row bands, the exchange buffers are
    stitched (what the all-gather does), post-processing runs on the full frame -> identical to a single run.  With spatial reuse
    each band also carries the row above and the row below it up to the tempDirectResv write (the halo launch of k_direct_stage)."""
    arrays = scenes.small_room()
    size = (256, 144)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    full, a, b = eid.Renderer(), eid.Renderer(), eid.Renderer()
    for r in (full, a, b):
        r.create(size, psc, acc)
        r.set_env_constant(common.ENV)
    a.set_band(0, 72)            # 72 / 2 = 36 quarter-res rows: the band edge lies in the MIDDLE of an 8x8 quarter-res tile row
    b.set_band(72, 144)
    info = psc.info()
    psc.update_camera(*size)
    exchange = [abi.BUF_THIS_GBUFFER, abi.BUF_MOTION, abi.BUF_DIRECT, abi.BUF_DENOISE_IND_A, abi.BUF_THIS_DIRECT_RESV,
                abi.BUF_THIS_INDIRECT_RESV]
    if not exchange_history:
        # what bench.py exchanges (static camera): no reservoir history crosses the ranks; with spatial reuse the halo rows keep their own
        # direct-reservoir history, so the rows next to a band edge still see the neighbour's reservoir of the previous frame
        exchange = [abi.BUF_THIS_GBUFFER, abi.BUF_DIRECT, abi.BUF_DENOISE_IND_A]
    cam = arrays.camera
    for f in range(4 if orbit else 3):
        if orbit:                 # moving camera: temporal reprojection crosses the band edge, so the reservoir history must be exchanged too
            ang = np.deg2rad(1.5 * f)
            e = np.array(cam["eye"], np.float64)
            psc.set_lookat((e[0] * np.cos(ang) - e[2] * np.sin(ang), e[1] + 0.15 * f, e[0] * np.sin(ang) + e[2] * np.cos(ang)), cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, ReSTIRState=restir)
        full.run(st, f)
        a.run_trace(st, f)
        b.run_trace(st, f)
        for which in exchange:
            ba, bb = a.read(which).view(np.uint8).copy(), b.read(which).view(np.uint8)
            _, off, n = b.band_range(which)
            ba[off:off + n] = bb[off:off + n]
            a.write(which, ba)
            b.write(which, ba)
        a.run_post(st, f)
        b.run_post(st, f)
        assert (a.stats().closestHitRays + b.stats().closestHitRays, a.stats().anyHitRays + b.stats().anyHitRays) == (
            full.stats().closestHitRays, full.stats().anyHitRays)          # (counters reach the host at the end of a frame) halo rows are not counted twice
        ref = common.snapshot(full)
        for r in (a, b):
            got = common.snapshot(r)
            for k in ref:
                if not exchange_history and k in ("motion", "direct_resv", "indirect_resv"):
                    continue                      # not gathered in this mode: every rank only holds its own rows (+ halo)
                assert got[k].tobytes() == ref[k].tobytes(), "band-sharded %s differs from the full-frame run (frame %d)" % (k, f)


def test_band_sharded_post_equals_full_frame():
    """Mode B: trace bands, exchange, then each rank denoises/composes only its band (levels evaluated on band + reach of the
    later levels); the final band rows equal the single-GPU frame bit for bit."""
    arrays = scenes.small_room()
    size = (256, 160)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    full = eid.Renderer()
    full.create(size, psc, acc)
    full.set_env_constant(common.ENV)
    bands = [(0, 40), (40, 104), (104, 160)]      # multiples of 8 rows; 40 / 2 = 20 and 104 / 2 = 52 are not multiples of 8 (mid-tile edges)
    ranks = []
    for y0, y1 in bands:
        r = eid.Renderer()
        r.create(size, psc, acc)
        r.set_env_constant(common.ENV)
        r.set_band(y0, y1)
        ranks.append(r)
    info = psc.info()
    psc.update_camera(*size)
    exchange = [abi.BUF_THIS_GBUFFER, abi.BUF_DIRECT, abi.BUF_DENOISE_IND_A]
    for f in range(3):
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f)
        full.run(st, f)
        for r in ranks:
            r.run_trace(st, f)
        for which in exchange:
            stitched = ranks[0].read(which).view(np.uint8).copy()
            for r in ranks[1:]:
                _, off, n = r.band_range(which)
                stitched[off:off + n] = r.read(which).view(np.uint8)[off:off + n]
            for r in ranks:
                r.write(which, stitched)
        for r in ranks:
            r.run_post_band(st, f)
        for which in (abi.BUF_DIRECT, abi.BUF_INDIRECT):
            want = full.read(which).view(np.uint8)
            for r in ranks:
                _, off, n = r.band_range(which)
                got = r.read(which).view(np.uint8)
                assert got[off:off + n].tobytes() == want[off:off + n].tobytes(), "mode B band rows differ (buffer %d, frame %d)" % (which, f)


def test_render_host_end_to_end_and_checkpoint():
    """eid_renderer_render_host (host buffers in, host images out) == run + read; history write-back resumes bit-exactly."""
    arrays = scenes.cornell_scene()
    size = (128, 72)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    r1, r2 = eid.Renderer(), eid.Renderer()
    for r in (r1, r2):
        r.create(size, psc, acc)
        r.set_env_constant(common.ENV)
    info = psc.info()
    psc.update_camera(*size)
    d = np.zeros((size[1], size[0], 4), np.float32)
    i = np.zeros_like(d)
    for f in range(2):
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f)
        r1.run(st, f)
        r2.render_host(psc.get_camera(), st, f, d.ctypes.data, i.ctypes.data)
        assert r1.read(abi.BUF_DIRECT).tobytes() == d.tobytes() and r1.read(abi.BUF_INDIRECT).tobytes() == i.tobytes()
    # pipelined host API: same images, delivered after wait_host; two buffer pairs alternate
    r4 = eid.Renderer()
    r4.create(size, psc, acc)
    r4.set_env_constant(common.ENV)
    bufs = [[np.zeros_like(d), np.zeros_like(d)] for _ in range(2)]
    cams = []
    psc2 = eid.Scene(0)
    psc2.load_arrays(arrays)
    psc2.update_camera(*size)
    for f in range(3):
        psc2.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f)
        r4.render_host_async(psc2.get_camera(), st, f, bufs[f & 1][0].ctypes.data, bufs[f & 1][1].ctypes.data)
    r4.wait_host()
    r5 = eid.Renderer()
    r5.create(size, psc, acc)
    r5.set_env_constant(common.ENV)
    psc3 = eid.Scene(0)
    psc3.load_arrays(arrays)
    psc3.update_camera(*size)
    for f in range(3):
        psc3.update_camera(*size)
        r5.render_host(psc3.get_camera(), common.frame_state(size[0], size[1], info, f), f, d.ctypes.data, i.ctypes.data)
        if f >= 1:
            assert bufs[f & 1][0].tobytes() == d.tobytes() and bufs[f & 1][1].tobytes() == i.tobytes(), "async host frame %d differs" % f
    # checkpoint = cross-frame state (G-buffer + reservoirs); restore into a fresh renderer and continue
    r3 = eid.Renderer()
    r3.create(size, psc, acc)
    r3.set_env_constant(common.ENV)
    psc.update_camera(*size)
    st = common.frame_state(size[0], size[1], info, 2)
    # after frame 1 (set 0): this* = [1]; frame 2 uses set 1 -> last* = [1]. Seed r3 so its "last" buffers hold r1's "this".
    r3.run(common.frame_state(size[0], size[1], info, 1), 1)          # selects the same ping-pong parity as r1
    for which in (abi.BUF_THIS_GBUFFER, abi.BUF_THIS_DIRECT_RESV, abi.BUF_THIS_INDIRECT_RESV):
        r3.write(which, r1.read(which))
    r1.run(st, 2)
    r3.run(st, 2)
    for k, which in common.BUFFERS:
        assert r1.read(which).tobytes() == r3.read(which).tobytes(), k


def test_golden_frames():
    """Committed oracle dumps (tests/golden/frames_*.npz, made by tests/golden/make_golden.py) == CUDA output."""
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    files = sorted(f for f in os.listdir(gdir) if f.startswith("frames_") and f.endswith(".npz"))
    assert files, "no golden frame fixtures"
    import make_golden_cfg as cfg
    for fn in files:
        z = np.load(os.path.join(gdir, fn))
        name = fn[len("frames_"):-4]
        maker, size, frames, over = cfg.CONFIGS[name][:4]
        arrays = maker()
        psc = eid.Scene(0)
        psc.load_arrays(arrays)
        acc = eid.AccelStructure()
        acc.create(psc)
        prr = eid.Renderer()
        prr.create(size, psc, acc)
        prr.set_env_constant(common.ENV)
        prr.set_strict_math(True)
        ss = cfg.sun_sky_of(cfg.CONFIGS[name])
        if ss is not None:
            prr.set_sun_and_sky(ss)
        psc.update_camera(*size)
        info = psc.info()
        for f in range(frames):
            psc.update_camera(*size)
            prr.run(common.frame_state(size[0], size[1], info, f, **over), f)
        got = common.snapshot(prr)
        if "display" in z.files:
            prr.run_output(abi.default_tonemapper())
            prr.sync()
            disp = prr.read(abi.BUF_DISPLAY_F32).reshape(size[1], size[0], 4)
            assert disp.view(np.uint32).tobytes() == z["display"].view(np.uint32).tobytes(), "golden %s: display pass differs" % name
        want = {k: z[k].view(got[k].dtype) if got[k].dtype.fields is None else np.frombuffer(z[k].tobytes(), got[k].dtype) for k in got}
        rep = common.compare_snapshots(got, want, "golden " + name)
        assert all(v == 0.0 for v in rep.values()), "golden %s: strict math must be bit-exact, got %s" % (name, rep)


def test_error_behaviour_gpu():
    psc = eid.Scene(0)
    with pytest.raises(eid.EidolaError):
        eid.AccelStructure().create(psc)            # no scene loaded
    psc.load_arrays(scenes.cube_scene())
    acc = eid.AccelStructure()
    acc.create(psc)
    r = eid.Renderer()
    r.create((64, 64), psc, acc)
    info = psc.info()
    psc.update_camera(64, 64)
    with pytest.raises(eid.EidolaError):
        r.run(common.frame_state(128, 64, info, 0), 0)                       # size beyond the allocation
    with pytest.raises(eid.EidolaError):
        r.run(common.frame_state(64, 64, info, 0, environmentProb=0.25), 0)  # needs an HDR map (sun & sky not implemented)
    with pytest.raises(eid.EidolaError):
        r.set_band(4, 64)                                                     # band edges must be multiples of 8
    r.run(common.frame_state(64, 64, info, 0), 0)                             # still usable after errors
    r.sync()


@pytest.mark.parametrize("mode", ["replicated", "sharded"])
def test_interleaved_stripes_equal_full_frame(mode):
    """Interleaved stripe ownership (load balance): 3 ranks x 2 exchange groups emulated on one device; the per-group
    in-place all-gather is played by read/stitch/write through eid_renderer_exchange_range."""
    from eidola_b200 import sharding
    arrays = scenes.small_room()
    w, h, world, groups = 192, 150, 3, 2
    srows, alloc_h = sharding.stripe_layout(h, world, groups)
    assert srows % 16 == 0 and alloc_h == srows * world * groups and alloc_h >= h
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    full = eid.Renderer()
    full.create((w, alloc_h), psc, acc)
    full.set_env_constant(common.ENV)
    ranks = []
    for k in range(world):
        r = eid.Renderer()
        r.create((w, alloc_h), psc, acc)
        r.set_env_constant(common.ENV)
        r.set_stripes(k, world, srows)
        assert r.exchange_groups() == groups
        ranks.append(r)
    info = psc.info()
    psc.update_camera(w, h)

    def all_gather(buffers):
        for which in buffers:
            stitched = ranks[0].read(which).view(np.uint8).copy()
            for r in ranks[1:]:
                mine = r.read(which).view(np.uint8)
                for g in range(groups):
                    _, off, n = r.exchange_range(which, g)
                    stitched[off:off + n] = mine[off:off + n]
            for r in ranks:
                r.write(which, stitched)

    for f in range(3):
        psc.update_camera(w, h)
        st = common.frame_state(w, h, info, f)
        full.run(st, f)
        for r in ranks:
            r.run_trace(st, f)
        all_gather([abi.BUF_THIS_GBUFFER, abi.BUF_DIRECT, abi.BUF_DENOISE_IND_A])
        if mode == "replicated":
            for r in ranks:
                r.run_post(st, f)
        else:
            for r in ranks:
                r.run_post_band(st, f)
            all_gather([abi.BUF_DIRECT, abi.BUF_INDIRECT])
        for which in (abi.BUF_DIRECT, abi.BUF_INDIRECT):
            want = full.read(which).view(np.uint8).reshape(alloc_h, -1)[:h]
            for r in ranks:
                got = r.read(which).view(np.uint8).reshape(alloc_h, -1)[:h]
                assert got.tobytes() == want.tobytes(), "%s stripes: buffer %d differs from the full-frame run (frame %d)" % (mode, which, f)
    # ray counters: the ranks together issue exactly the rays of the single-GPU frame
    tot = [sum(getattr(r.stats(), k) for r in ranks) for k in ("closestHitRays", "anyHitRays")]
    assert tot == [full.stats().closestHitRays, full.stats().anyHitRays]


def test_group_of_one_equals_renderer_and_delivers_bands():
    """eid_group with world = 1 (no NCCL involved): the group schedule == eid_renderer_run, and eid_group_render_host_async delivers the
    band (here: the whole frame) into host images like eid_renderer_render_host does.  The N > 1 collectives are exercised by bench.py
    under torchrun (image_crc32 of the N-GPU frame == the 1-GPU frame) and, stage by stage, by the band tests above."""
    arrays = scenes.cornell_scene()
    size = (160, 96)
    psc = eid.Scene(0); psc.load_arrays(arrays)
    acc = eid.AccelStructure(); acc.create(psc)
    r1, r2 = eid.Renderer(), eid.Renderer()
    for r in (r1, r2):
        r.create(size, psc, acc); r.set_env_constant(common.ENV)
    assert eid.Group.layout(1080, 8, 7) == (952, 1088, 1088) and eid.Group.layout(1080, 1) == (0, 1080, 1080)
    assert eid.Group.layout(size[1], 2, 1) == (48, 96, 96)
    g = eid.Group()
    g.create(r2, 0, 1)
    g.set_mode(True, 2, True)
    info = psc.info()
    psc.update_camera(*size)
    d = [np.zeros((size[1], size[0], 4), np.float32) for _ in range(2)]
    i = [np.zeros_like(d[0]) for _ in range(2)]
    for f in range(4):
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f)
        r1.run(st, f)
        if f < 2:
            g.run(st, f); g.sync()
        else:
            g.render_host_async(psc.get_camera(), st, f, d[f & 1].ctypes.data, i[f & 1].ctypes.data); g.wait_host()
            assert r1.read(abi.BUF_DIRECT).tobytes() == d[f & 1].tobytes() and r1.read(abi.BUF_INDIRECT).tobytes() == i[f & 1].tobytes()
        for k, which in common.BUFFERS:
            assert r1.read(which).tobytes() == r2.read(which).tobytes(), "group frame %d: %s differs" % (f, k)
    gi = g.info()
    assert (gi.rank, gi.world, gi.y0, gi.y1, gi.collectives) == (0, 1, 0, size[1], 0)
    with pytest.raises(eid.EidolaError):
        eid.Group().create(r1, 0, 2, None)                                    # world > 1 needs the unique id
    with pytest.raises(eid.EidolaError):
        eid.Group().create(r1, 3, 2, bytes(128))


def test_environment_reload_keeps_the_renderers_map_alive():
    """HdrSampling::loadEnvironment on a live object (sample_example.cpp:97-106) destroys and re-creates the map while a renderer still
    samples the previous one: the library keeps that map alive until the renderer is given the new one (no use-after-free)."""
    import oracle_lib as ol
    arrays = scenes.cube_scene()
    size = (96, 64)
    sky1, sky2 = scenes.synthetic_sky(), scenes.synthetic_sky(sun=False)
    osc, orr, psc, acc, prr = common.make_pair(arrays, size, env_img=sky1)
    penv = prr._env
    info = psc.info()
    over1 = common.env_state_overrides(penv.get_integral())
    for s in (osc, psc):
        s.update_camera(*size)

    def frame(f, over):
        for s in (osc, psc):
            s.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, **over)
        orr.run(st, f); prr.run(st, f); prr.sync()
        rep = common.compare_snapshots(common.snapshot(prr), common.snapshot(orr), "env reload frame %d" % f)
        assert all(v == 0.0 for v in rep.values()), rep
    frame(0, over1)
    penv.set_pixels(sky2)                       # destroy + create on the same HdrSampling object; the renderer was NOT told
    frame(1, over1)                             # still the first map, bit for bit
    oenv2 = ol.OracleEnv(sky2)
    orr.set_env(oenv2)
    prr.set_env(penv)                           # now the renderer switches (and the first map is freed)
    frame(2, common.env_state_overrides(penv.get_integral()))
    prr.set_env(None)
    penv.destroy()


@pytest.mark.parametrize("restir", [abi.eTemporal, abi.eSpatiotemporal])
def test_frames_in_flight_are_bit_identical(restir):
    """eid_renderer_set_pipeline(2): direct_stage of frame f + 1 overlaps indirect_stage / denoise / compose of frame f (its own stream, per-parity
    direct image / K2 gather buffers / counters).  A burst of frames enqueued without any host synchronisation — moving camera, so every temporal
    lookup really reads the previous frame — must leave exactly the buffers the strictly ordered renderer leaves, frame after frame."""
    arrays = scenes.small_room()
    size = (448, 256)
    psc = eid.Scene(0); psc.load_arrays(arrays)
    acc = eid.AccelStructure(); acc.create(psc)
    strict, piped, hosted = eid.Renderer(), eid.Renderer(), eid.Renderer()
    for r in (strict, piped, hosted):
        r.create(size, psc, acc); r.set_env_constant(common.ENV)
    piped.set_pipeline(2); hosted.set_pipeline(2)
    info = psc.info()
    cam = arrays.camera
    psc.update_camera(*size)
    bufs = [[np.zeros((size[1], size[0], 4), np.float32) for _ in range(2)] for _ in range(2)]
    f = 0
    for burst in (1, 5, 2, 7):
        cams, states = [], []
        for _ in range(burst):
            a = np.deg2rad(0.8 * f)
            e = np.array(cam["eye"], np.float64)
            psc.set_lookat((e[0] * np.cos(a) - e[2] * np.sin(a), e[1], e[0] * np.sin(a) + e[2] * np.cos(a)), cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
            psc.update_camera(*size)
            st = common.frame_state(size[0], size[1], info, f, ReSTIRState=restir, maxDepth=3)
            strict.run(st, f); strict.sync()
            piped.run(st, f)                                        # no sync inside the burst
            hosted.render_host_async(psc.get_camera(), st, f, bufs[f & 1][0].ctypes.data, bufs[f & 1][1].ctypes.data)
            f += 1
        piped.sync(); hosted.wait_host(); hosted.sync()
        want = common.snapshot(strict)
        for name, r in (("pipelined", piped), ("pipelined host", hosted)):
            got = common.snapshot(r)
            for k in want:
                assert got[k].tobytes() == want[k].tobytes(), "%s: %s differs after frame %d" % (name, k, f - 1)
            s0, s1 = strict.stats(), r.stats()
            assert (s0.closestHitRays, s0.anyHitRays, s0.primaryHits) == (s1.closestHitRays, s1.anyHitRays, s1.primaryHits)
            assert (s0.totalClosestHitRays, s0.totalAnyHitRays) == (s1.totalClosestHitRays, s1.totalAnyHitRays)
        k = (f - 1) & 1
        assert strict.read(abi.BUF_DIRECT).tobytes() == bufs[k][0].tobytes() and strict.read(abi.BUF_INDIRECT).tobytes() == bufs[k][1].tobytes()
    # strictly ordered entry points mixed in: run_trace + run_post on the pipelined renderer, then pipelined frames again
    for mixed in range(3):
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, ReSTIRState=restir, maxDepth=3)
        strict.run(st, f)
        if mixed == 1:
            piped.run_trace(st, f); piped.run_post(st, f)
        else:
            piped.run(st, f)
        f += 1
    piped.sync(); strict.sync()
    want, got = common.snapshot(strict), common.snapshot(piped)
    for k in want:
        assert got[k].tobytes() == want[k].tobytes(), "mixed strict / pipelined entry points: %s differs" % k


@pytest.mark.parametrize("variant", [abi.VARIANT_DIRECT_BILATERAL, abi.VARIANT_INDIRECT_BILATERAL, abi.VARIANT_FETCH_4_SUBPIXELS,
                                     abi.VARIANT_DIRECT_BILATERAL | abi.VARIANT_INDIRECT_BILATERAL | abi.VARIANT_FETCH_4_SUBPIXELS,
                                     abi.VARIANT_DIRECT_SPLIT, abi.VARIANT_DIRECT_SPLIT | abi.VARIANT_INDIRECT_BILATERAL | abi.VARIANT_FETCH_4_SUBPIXELS])
@pytest.mark.parametrize("wavefront", [True, False])
def test_reference_compile_time_variants(variant, wavefront):
    """Scope row (f.4): the reference's dormant variants as run-time switches (eid_renderer_set_variant) — the two-kernel direct stage
    direct_gen.comp + direct_reuse.comp (built by the reference, renderer.cpp:129-132, never dispatched), the bilateral
    denoisers (DENOISER_DIRECT_BILATERAL: direct_stage.comp:284-288, denoise_direct.comp:73-137, renderer.cpp:186-188;
    DENOISER_INDIRECT_BILATERAL: denoise_indirect.comp:77-130) and FETCH_GEOM_CHECK_4_SUBPIXELS (pathtrace.glsl:314-358, one more RNG
    draw per quarter-res pixel) — bit-identical to the oracle with strict math, both K2 forms, moving camera, sky pixels included."""
    for maker, size, kw in ((scenes.small_room, (200, 120), dict(orbit=True)), (scenes.cube_scene, (96, 64), dict()), (scenes.cornell_scene, (130, 70), dict(denoise=0))):
        arrays = maker()
        osc, orr, psc, acc, prr = common.make_pair(arrays, size)
        prr.set_wavefront(wavefront)
        orr.set_variant(variant); prr.set_variant(variant)
        for s in (osc, psc):
            s.update_camera(*size)
        info = psc.info()
        cam = arrays.camera
        over = {k: v for k, v in kw.items() if k != "orbit"}
        for f in range(3):
            if kw.get("orbit"):
                a = np.deg2rad(0.7 * f)
                e = np.array(cam["eye"], np.float64)
                for s in (osc, psc):
                    s.set_lookat((e[0] * np.cos(a) - e[2] * np.sin(a), e[1], e[0] * np.sin(a) + e[2] * np.cos(a)), cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
            for s in (osc, psc):
                s.update_camera(*size)
            st = common.frame_state(size[0], size[1], info, f, maxDepth=3, **over)
            orr.run(st, f); prr.run(st, f); prr.sync()
            got, want = common.snapshot(prr), common.snapshot(orr)
            rep = common.compare_snapshots(got, want, "variant %d %s frame %d" % (variant, maker.__name__, f))
            assert all(v == 0.0 for v in rep.values()), rep
            assert prr.read(abi.BUF_DENOISE_DIR_A).tobytes() == orr.read(abi.BUF_DENOISE_DIR_A).tobytes()
            so, sp = orr.stats(), prr.stats()
            assert (so.closestHitRays, so.anyHitRays) == (sp.closestHitRays, sp.anyHitRays)
    # default numerics: within the 1e-3 contract
    arrays = scenes.small_room()
    osc, orr, psc, acc, prr = common.make_pair(arrays, (200, 120), strict=False)
    orr.set_variant(variant); prr.set_variant(variant)
    for s in (osc, psc):
        s.update_camera(200, 120)
    for f in range(2):
        for s in (osc, psc):
            s.update_camera(200, 120)
        st = common.frame_state(200, 120, psc.info(), f, maxDepth=3)
        orr.run(st, f); prr.run(st, f); prr.sync()
        rep = common.compare_snapshots(common.snapshot(prr), common.snapshot(orr), "variant %d fast frame %d" % (variant, f))
        assert max(rep.values()) <= 1e-3
    with pytest.raises(eid.EidolaError):
        prr.set_variant(16)
