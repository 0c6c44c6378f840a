import sys, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from importlib import import_module
sc = import_module("cis-565-final-vr-raytracer_b200.scenes")
q = int(sys.argv[1]) if len(sys.argv) > 1 else 707
a = sc.heightfield_room(quads=q)
out = []
for n in a.nodes:
    pm = a.prim_meshes[n["primMesh"]]
    idx = a.indices[pm["firstIndex"]:pm["firstIndex"] + pm["indexCount"]].astype(np.int64) + pm["vertexOffset"]
    out.append(a.positions[idx].reshape(-1, 9))
t = np.concatenate(out).astype(np.float32)
print(t.shape)
t.tofile("tris_%d.bin" % q)
