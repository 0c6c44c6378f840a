"""K1 / K2 / post times of every 1/8 band of the C3 frame on one GPU (what each rank of an 8-GPU run renders): load imbalance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eidola_b200 as eid
import bench
W, H = 1920, 1080
arrays = bench.scene_arrays()
scene = eid.Scene(0); scene.load_arrays(arrays)
accel = eid.AccelStructure(); accel.create(scene)
info = scene.info()
world = 8
_, _, alloc = eid.Group.layout(H, world)
tot = np.zeros(5)
for rank in range(world):
    y0, y1, _ = eid.Group.layout(H, world, rank)
    rr = eid.Renderer(); rr.create((W, alloc), scene, accel); rr.set_env_constant(bench.ENV); rr.set_overlap(False)
    rr.set_band(y0, y1)
    scene.update_camera(W, H)
    acc = np.zeros(5); n = 0
    for f in range(8):
        scene.update_camera(W, H)
        st = bench.frame_state(info, f, W, H)
        rr.set_profiling(1)
        rr.run_trace(st, f); rr.run_post_band(st, f)
        s = rr.stats()
        if f >= 3: acc += np.array(s.kernelMs[:]); n += 1
    acc /= n; tot += acc
    print("band %d rows %4d-%4d  K1 %.3f K2 %.3f K3 %.3f K4 %.3f K5 %.3f  | K1+K2 %.3f" % (rank, y0, y1, *acc, acc[0] + acc[1]))
print("sum over bands: K1 %.3f K2 %.3f K3 %.3f K4 %.3f K5 %.3f" % tuple(tot))
