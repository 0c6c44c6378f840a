"""Two-level acceleration structure (BLAS per prim mesh + TLAS over the instances, AccelStructure::create, accelstruct.cpp:55-162) against the
flat world-space BVH that every other test pins to the oracle and to brute force: same hits bit for bit (closest hit with barycentrics
and ids, occlusion), same frames bit for bit (every buffer), for instanced / mirrored / far-from-origin / alpha-tested / textured scenes."""
import numpy as np
import pytest

import eidola_b200 as eid
from eidola_b200 import abi, scenes

import common

pytestmark = pytest.mark.gpu


def _rays(arrays, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    cam = arrays.camera
    # world-space scene bounds are not known without the node transforms: shoot from around the camera and its target
    c = np.array(cam["center"], np.float64)
    e = np.array(cam["eye"], np.float64)
    o = c + (rng.random((n, 3)) - 0.5) * 2.0 * np.linalg.norm(e - c)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = np.where(rng.random(n) < 0.5, 1e30, rng.random(n) * 6.0)
    rays[:, 4:7] = d
    rays[: n // 8, 4:7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 8)] * rng.choice([-1.0, 1.0], n // 8)[:, None]   # axis-parallel rays
    return rays


SCENES = [("instanced", scenes.instanced_scene, {}), ("grid", scenes.instanced_grid, {}),
          ("grid_far", scenes.instanced_grid, dict(offset=(1500.0, -300.0, 2500.0))), ("alpha", scenes.alpha_scene, {}),
          ("textured", scenes.textured_scene, {}), ("room", scenes.small_room, {})]


@pytest.mark.parametrize("name,maker,kw", SCENES, ids=[s[0] for s in SCENES])
def test_two_level_hits_equal_flat(name, maker, kw):
    arrays = maker(**kw)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    flat, two = eid.AccelStructure(), eid.AccelStructure()
    flat.create(psc, abi.ACCEL_FLAT)
    two.create(psc, abi.ACCEL_TWO_LEVEL)
    fi, ti = flat.info(), two.info()
    assert fi.twoLevel == 0 and ti.twoLevel == 1 and ti.instanceCount == len(arrays.nodes) and ti.triangleCount <= fi.triangleCount
    rays = _rays(arrays, 200000, 5)
    a, b = flat.trace(rays), two.trace(rays)
    assert (a["hitT"] < 1e27).sum() > 1000, "the rays must hit something for this to mean anything"
    assert a.tobytes() == b.tobytes(), "%s: %d of %d closest hits differ between the flat and the two-level structure" % (
        name, int((a.view(np.uint8).reshape(len(a), -1) != b.view(np.uint8).reshape(len(b), -1)).any(axis=1).sum()), len(a))
    oa, ob = flat.trace(rays, any_hit=True)["hitT"], two.trace(rays, any_hit=True)["hitT"]
    assert np.array_equal(oa, ob), "%s: occlusion differs for %d rays" % (name, int((oa != ob).sum()))


@pytest.mark.parametrize("name,maker,kw", SCENES[:5], ids=[s[0] for s in SCENES[:5]])
def test_two_level_frames_equal_flat(name, maker, kw):
    arrays = maker(**kw)
    size = (192, 112)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    rr = []
    for mode in (abi.ACCEL_FLAT, abi.ACCEL_TWO_LEVEL):
        acc = eid.AccelStructure()
        acc.create(psc, mode)
        r = eid.Renderer()
        r.create(size, psc, acc)
        r.set_env_constant(common.ENV)
        r.set_strict_math(True)
        rr.append((acc, r))
    info = psc.info()
    psc.update_camera(*size)
    cam = arrays.camera
    for f in range(3):
        ang = np.deg2rad(1.0 * f)
        e = np.array(cam["eye"], np.float64) - np.array(cam["center"], np.float64)
        c = np.array(cam["center"], np.float64)
        psc.set_lookat((c[0] + e[0] * np.cos(ang) - e[2] * np.sin(ang), c[1] + e[1], c[2] + e[0] * np.sin(ang) + e[2] * np.cos(ang)), cam["center"], cam["up"], np.rad2deg(cam["yfov"]))
        psc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, maxDepth=3)
        snaps = []
        for acc, r in rr:
            r.run(st, f)
            r.sync()
            snaps.append((common.snapshot(r), r.stats()))
        (sa, ta), (sb, tb) = snaps
        for k in sa:
            assert sa[k].tobytes() == sb[k].tobytes(), "%s frame %d: %s differs between the flat and the two-level structure" % (name, f, k)
        assert (ta.closestHitRays, ta.anyHitRays, ta.primaryHits) == (tb.closestHitRays, tb.anyHitRays, tb.primaryHits)


def test_auto_mode_and_memory():
    """eid_accel_build picks the two-level form once instancing would at least double the flat triangle list; its memory does not grow with
    the instance count."""
    psc = eid.Scene(0)
    psc.load_arrays(scenes.instanced_grid(n=9))
    auto, flat = eid.AccelStructure(), eid.AccelStructure()
    auto.create(psc)
    flat.create(psc, abi.ACCEL_FLAT)
    ai, fi = auto.info(), flat.info()
    assert ai.twoLevel == 1 and ai.blasCount == 3 and ai.instanceCount == 83
    assert ai.triangleCount == 10 + 264 + 2 and fi.triangleCount == 10 + 81 * 264 + 2
    assert ai.triBytes + ai.nodeBytes < (fi.triBytes + fi.nodeBytes) / 10
    psc2 = eid.Scene(0)
    psc2.load_arrays(scenes.small_room())
    a2 = eid.AccelStructure()
    a2.create(psc2)
    assert a2.info().twoLevel == 0                    # nothing is instanced: the flat tree (faster walk) stays
