"""K1 / K2 times on row bands of the C3 frame (what one rank of an N-GPU run renders), for both forms of indirect_stage."""
import numpy as np
import eidola_b200 as eid
from eidola_b200 import abi
import bench
W, H = 1920, 1080
arrays = bench.scene_arrays()
scene = eid.Scene(0); scene.load_arrays(arrays)
accel = eid.AccelStructure(); accel.create(scene)
info = scene.info()
for label, band in (("full", None), ("1/2: rows 0-544", (0, 544)), ("1/2: rows 544-1088", (544, 1088)), ("1/4: rows 272-544", (272, 544)), ("1/4: rows 816-1088", (816, 1088)),
                    ("1/8: rows 0-144", (0, 144)), ("1/8: rows 432-576", (432, 576)), ("1/8: rows 1008-1152", (1008, 1152))):
    for form in (1, 0):
        rr = eid.Renderer(); rr.create((W, 1152), scene, accel); rr.set_env_constant(bench.ENV); rr.set_wavefront(form); rr.set_overlap(False)
        if band: rr.set_band(*band)
        scene.update_camera(W, H)
        acc = np.zeros(5); n = 0
        for f in range(8):
            scene.update_camera(W, H)
            st = bench.frame_state(info, f, W, H)
            rr.set_profiling(1)
            rr.run(st, f)
            s = rr.stats()
            if f >= 3: acc += np.array(s.kernelMs[:]); n += 1
        print("%-22s %-9s K1 %.3f K2 %.3f" % (label, "wavefront" if form else "mega", acc[0]/n, acc[1]/n))
