"""Shared helpers for the parity tests: scene setup on either side + buffer comparison."""
import numpy as np

import eidola_b200 as eid
from eidola_b200 import abi

ENV = (0.25, 0.25, 0.25)   # constant environment of SURVEY.md §8(d): integral = pi


def frame_state(w, h, info, frame, **over):
    """RtxState of SURVEY.md §8(d): reference defaults, environmentProb = 0, time = 1000 + 16*frame."""
    kw = dict(environmentProb=0.0, time=1000 + 16 * frame, fireflyClampThreshold=float(np.float32(4 * np.pi)),
              envMapLuminIntegInv=float(np.float32(1 / np.pi)),
              lightLuminIntegInv=float(np.float32(1.0) / (np.float32(info.trigLightWeight) + np.float32(info.puncLightWeight))))
    kw.update(over)
    return abi.default_rtx_state(w, h, **kw)


BUFFERS = [("gbuffer", abi.BUF_THIS_GBUFFER), ("motion", abi.BUF_MOTION), ("direct_resv", abi.BUF_THIS_DIRECT_RESV),
           ("indirect_resv", abi.BUF_THIS_INDIRECT_RESV), ("direct", abi.BUF_DIRECT), ("indirect", abi.BUF_INDIRECT),
           ("ind_tmp_a", abi.BUF_DENOISE_IND_A), ("ind_tmp_b", abi.BUF_DENOISE_IND_B)]


def snapshot(renderer):
    return {name: renderer.read(which).copy() for name, which in BUFFERS}


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def same_bits_or_both_nan(a, b):
    """Bit-identical floats, where a NaN only has to meet a NaN: the payload / sign of an invalid operation's NaN (0/0 in post.frag's
    toneExposure on a black pixel) is the processor's choice — 0xFFC00000 on x86, 0x7FFFFFFF on the GPU."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    d[both_nan] = 0.0
    return float(np.nanmax(d)) if d.size else 0.0


def compare_snapshots(got, want, tag, tol=1e-3):
    """Contract of BASELINE.json north_star: integer outputs (G-buffer words, motion indices, reservoir `num`)
    and the reservoir picks (Li, wi, dist / the GI sample) bit-exact; weights and radiance within `tol`
    relative (absolute floor 1e-6).  Returns a dict of the worst deviations (0.0 everywhere == bit-exact)."""
    report = {}
    assert np.array_equal(got["gbuffer"], want["gbuffer"]), "%s: G-buffer words differ in %d texels" % (
        tag, int((got["gbuffer"].reshape(-1, 4) != want["gbuffer"].reshape(-1, 4)).any(axis=1).sum()))
    assert np.array_equal(got["motion"], want["motion"]), "%s: motion indices differ" % tag
    g, w = got["direct_resv"], want["direct_resv"]
    assert np.array_equal(g["num"], w["num"]), "%s: direct reservoir num differs" % tag
    assert bits(g["lightSample"]).tobytes() == bits(w["lightSample"]).tobytes(), "%s: direct reservoir picks (Li, wi, dist) differ" % tag
    report["direct_resv.weight"] = rel_err(g["weight"], w["weight"])
    g, w = got["indirect_resv"], want["indirect_resv"]
    assert np.array_equal(g["num"], w["num"]), "%s: indirect reservoir num differs" % tag
    gs, ws = g["giSample"], w["giSample"]
    for f in ("xv", "nv", "xs", "ns"):
        assert bits(gs[f]).tobytes() == bits(ws[f]).tobytes(), "%s: GI sample %s differs" % (tag, f)
    report["indirect_resv.L"] = rel_err(gs["L"], ws["L"])
    report["indirect_resv.weight"] = rel_err(g["weight"], w["weight"])
    for name in ("direct", "indirect", "ind_tmp_a", "ind_tmp_b"):
        report[name] = rel_err(got[name], want[name])
    for k, v in report.items():
        assert v <= tol, "%s: %s deviates by %.3g (tolerance %.1e relative)" % (tag, k, v, tol)
    return report


def env_state_overrides(integral, prob=0.25):
    """RtxState fields SampleExample::loadEnvironmentHdr derives from the HDR map (sample_example.cpp:104-105) + the default
    environmentProb (sample_example.hpp:161)."""
    return dict(environmentProb=prob, fireflyClampThreshold=float(np.float32(integral) * np.float32(4.0)),
                envMapLuminIntegInv=float(np.float32(1.0) / np.float32(integral)))


def make_pair(arrays, size, use_bvh=True, strict=True, env_img=None):
    """(oracle scene, oracle renderer, product scene, product accel, product renderer) for one scene."""
    import oracle_lib as ol
    w, h = size
    osc = ol.OracleScene(use_bvh=use_bvh)
    osc.load_arrays(arrays)
    orr = ol.OracleRenderer(osc, size)
    orr.set_env_constant(ENV)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    prr = eid.Renderer()
    prr.create(size, psc, acc)
    prr.set_env_constant(ENV)
    prr.set_strict_math(strict)   # strict: deterministic exp in the denoiser -> every buffer bit-identical to the oracle
    if env_img is not None:
        oenv = ol.OracleEnv(env_img)
        penv = eid.HdrSampling(0)
        penv.set_pixels(env_img)
        assert oenv.accel().tobytes() == penv.accel().tobytes() and oenv.get_integral() == penv.get_integral()
        orr.set_env(oenv)
        prr.set_env(penv)
    return osc, orr, psc, acc, prr
