"""Longest rays of the C3 frame (profiling level 2): worst K1 thread, longest queued closest-hit / any-hit ray."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import eidola_b200 as eid
from eidola_b200 import abi, scenes
import common
arrays = scenes.heightfield_room()
sc = eid.Scene(0); sc.load_arrays(arrays)
acc = eid.AccelStructure(); acc.create(sc)
w, h = 1920, 1080
r = eid.Renderer(); r.create((w, h), sc, acc); r.set_env_constant(common.ENV); r.set_profiling(2)
info = sc.info(); sc.update_camera(w, h)
for f in range(4):
    st = common.frame_state(w, h, info, f, maxDepth=3)
    r.run(st, f); r.sync()
    s = r.stats()
    print("frame", f, "closest", s.closestHitRays, "any", s.anyHitRays, "primaryHits", s.primaryHits, "nodes/ray %.2f" % (s.nodeVisits / (s.closestHitRays + s.anyHitRays)),
          "worst K1 thread", s.maxNodeVisitsPerThread, "longest queued closest / any ray", s.maxNodeVisitsPerQueuedRay[0], s.maxNodeVisitsPerQueuedRay[1])
