"""Build quality of the acceleration structure (the reference builds with VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR,
accelstruct.cpp:125-126,161): EID_ACCEL_FAST_TRACE (binned SAH on the host threads, the default) against EID_ACCEL_FAST_BUILD (Morton LBVH on
the GPU, what every round-1 test ran on).  The tree decides how many nodes a ray visits, never what it hits: same hits bit for bit, same
frames bit for bit, flat and two-level forms; and the SAH tree visits fewer nodes on the C3 kind of scene."""
import numpy as np
import pytest

import eidola_b200 as eid
from eidola_b200 import abi, scenes

import common
from test_gpu_two_level import SCENES, _rays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,maker,kw", SCENES, ids=[s[0] for s in SCENES])
@pytest.mark.parametrize("form", [abi.ACCEL_FLAT, abi.ACCEL_TWO_LEVEL], ids=["flat", "two_level"])
def test_fast_trace_hits_equal_fast_build(name, maker, kw, form):
    arrays = maker(**kw)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    lbvh, sah = eid.AccelStructure(), eid.AccelStructure()
    lbvh.create(psc, form | abi.ACCEL_FAST_BUILD)
    sah.create(psc, form | abi.ACCEL_FAST_TRACE)
    li, si = lbvh.info(), sah.info()
    assert li.fastTrace == 0 and si.fastTrace == 1 and li.triangleCount == si.triangleCount
    rays = _rays(arrays, 200000, 11)
    a, b = lbvh.trace(rays), sah.trace(rays)
    assert (a["hitT"] < 1e27).sum() > 1000
    assert a.tobytes() == b.tobytes(), "%s: %d of %d closest hits differ between the Morton and the SAH tree" % (
        name, int((a.view(np.uint8).reshape(len(a), -1) != b.view(np.uint8).reshape(len(b), -1)).any(axis=1).sum()), len(a))
    oa, ob = lbvh.trace(rays, any_hit=True)["hitT"], sah.trace(rays, any_hit=True)["hitT"]
    assert np.array_equal(oa, ob), "%s: occlusion differs for %d rays" % (name, int((oa != ob).sum()))


def test_default_build_is_fast_trace_and_flags_are_checked():
    psc = eid.Scene(0)
    psc.load_arrays(scenes.small_room())
    acc = eid.AccelStructure()
    acc.create(psc)
    assert acc.info().fastTrace == 1
    with pytest.raises(eid.EidolaError):
        eid.AccelStructure().create(psc, abi.ACCEL_FAST_TRACE | abi.ACCEL_FAST_BUILD)


@pytest.mark.parametrize("name,maker,kw", [SCENES[0], SCENES[3], SCENES[5]], ids=["instanced", "alpha", "room"])
def test_fast_trace_frames_equal_fast_build(name, maker, kw):
    arrays = maker(**kw)
    size = (192, 112)
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    rr = []
    for mode in (abi.ACCEL_FLAT | abi.ACCEL_FAST_BUILD, abi.ACCEL_FLAT | abi.ACCEL_FAST_TRACE):
        acc = eid.AccelStructure()
        acc.create(psc, mode)
        r = eid.Renderer()
        r.create(size, psc, acc)
        r.set_env_constant(common.ENV)
        r.set_strict_math(True)
        r.set_profiling(2)                      # level 2: the kernels count node visits
        rr.append((acc, r))
    info = psc.info()
    psc.update_camera(*size)
    for f in range(3):
        st = common.frame_state(size[0], size[1], info, f, maxDepth=3)
        snaps = []
        for acc, r in rr:
            r.run(st, f)
            r.sync()
            snaps.append((common.snapshot(r), r.stats()))
        (sa, ta), (sb, tb) = snaps
        for k in sa:
            assert sa[k].tobytes() == sb[k].tobytes(), "%s frame %d: %s differs between the Morton and the SAH tree" % (name, f, k)
        assert (ta.closestHitRays, ta.anyHitRays, ta.primaryHits) == (tb.closestHitRays, tb.anyHitRays, tb.primaryHits)
    if name == "room":     # a height field: the case the Morton build handles worst (its height bits split patches into overlapping boxes)
        assert tb.nodeVisits < ta.nodeVisits, "SAH tree visits %d nodes, Morton tree %d" % (tb.nodeVisits, ta.nodeVisits)
