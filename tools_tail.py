"""Is K2 straggler-bound?  Times K2 alone on row subsets and reports the worst per-thread node count."""
import numpy as np
import eidola_b200 as eid
from eidola_b200 import abi
import bench
W, H = 1920, 1080
arrays = bench.scene_arrays()
scene = eid.Scene(0); scene.load_arrays(arrays)
accel = eid.AccelStructure(); accel.create(scene)
info = scene.info()
for label, band in (("full", None), ("rows 0-544", (0, 544)), ("rows 544-1088", (544, 1088)), ("rows 0-272", (0, 272)), ("rows 272-544", (272, 544)), ("rows 816-1088", (816, 1088))):
    rr = eid.Renderer(); rr.create((W, 1088), scene, accel); rr.set_env_constant(bench.ENV)
    if band: rr.set_band(*band)
    scene.update_camera(W, H)
    acc = np.zeros(5); n = 0
    for f in range(8):
        scene.update_camera(W, H)
        st = bench.frame_state(info, f, W, H)
        rr.set_profiling(2 if f == 7 else 1)
        rr.run(st, f)
        s = rr.stats()
        if 2 <= f < 7: acc += np.array(s.kernelMs[:]); n += 1
    print("%-16s K1 %.3f K2 %.3f | max node visits by one thread %d, mean nodes/ray %.1f" % (label, acc[0]/n, acc[1]/n, s.maxNodeVisitsPerThread, s.nodeVisits/max(1, s.closestHitRays+s.anyHitRays)))
