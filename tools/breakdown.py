"""Where does the time of K1/K2 go?  Runs the C3 frame with RtxState variants and prints the per-stage CUDA-event times."""
import sys
import numpy as np
import torch
import eidola_b200 as eid
from eidola_b200 import abi, scenes
import bench

W, H = 1920, 1080
arrays = bench.scene_arrays()
scene = eid.Scene(0); scene.load_arrays(arrays)
accel = eid.AccelStructure(); accel.create(scene)
rr = eid.Renderer(); rr.create((W, H), scene, accel); rr.set_env_constant(bench.ENV)
info = scene.info()
scene.update_camera(W, H)
variants = [("default", {}), ("debug_normal(primary only)", dict(debugging_mode=abi.eNormal)), ("eNone(1 light+shadow)", dict(ReSTIRState=abi.eNone)),
            ("RIS M=1", dict(ReSTIRState=abi.eRIS, RISSampleNum=1)), ("RIS M=4", dict(ReSTIRState=abi.eRIS)), ("RIS M=8", dict(ReSTIRState=abi.eRIS, RISSampleNum=8)),
            ("maxDepth=1", dict(maxDepth=1)), ("maxDepth=2", dict(maxDepth=2)), ("maxDepth=4", dict(maxDepth=4)), ("MIS=0", dict(MIS=0))]
frame = 0
for name, over in variants:
    rr.set_profiling(1)
    acc = np.zeros(5); n = 0
    for k in range(6):
        scene.update_camera(W, H)
        st = bench.frame_state(info, frame, W, H)
        for a, b in over.items():
            setattr(st, a, b)
        rr.run(st, frame); frame += 1
        s = rr.stats()
        if k >= 2:
            acc += np.array(s.kernelMs[:]); n += 1
    print("%-28s K1 %.3f K2 %.3f K3 %.3f K4 %.3f | rays closest %d any %d nodes/ray %.1f tris/ray %.1f" % (
        name, acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, s.closestHitRays, s.anyHitRays,
        s.nodeVisits / max(1, s.closestHitRays + s.anyHitRays), s.triangleTests / max(1, s.closestHitRays + s.anyHitRays)))
