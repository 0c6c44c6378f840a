"""debug helper: run a pipeline config and print, per rank / buffer / frame, the rows that differ from the single-GPU frame"""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import test_gpu_pipeline as t
import pipeline_worker as pw

world = int(sys.argv[1]); cfg = dict(scene="small_room", size=[256, 144], stages=[0, 0, 0], history=2, frames=3)
cfg.update(json.loads(sys.argv[2]) if len(sys.argv) > 2 else {})
w, h = cfg["size"]
want, rays = t.single_gpu_frames(cfg)
ranks = t.run_ranks(world, cfg)
owned, _ = pw.owned_tables()
print("rays", rays, [(int(r["meta"][3]), int(r["meta"][4])) for r in ranks], "memops", [int(r["meta"][7]) for r in ranks])
for rank, r in enumerate(ranks):
    role, y0, y1 = int(r["meta"][0]), int(r["meta"][1]), int(r["meta"][2])
    for key in sorted(k for k in r if k != "meta" and not k.startswith("rows_")):
        name, f = key.rsplit("_", 1)
        which, row_bytes, half = owned[name]
        a, b = (int(v) for v in r["rows_" + name])
        rb = row_bytes(w)
        ref = want[int(f)][name][a * rb:b * rb].reshape(b - a, rb)
        got = r[key].reshape(b - a, rb)
        bad = np.nonzero((got != ref).any(axis=1))[0]
        if bad.size and name in ("direct", "indirect"):
            g4 = got.view(np.float32).reshape(b - a, w, 4); r4 = ref.view(np.float32).reshape(b - a, w, 4)
            dpx = np.nonzero((g4 != r4).any(axis=2))
            print("   %d pixels differ; examples (y, x, got, want):" % dpx[0].size)
            for k in range(0, dpx[0].size, max(1, dpx[0].size // 8)):
                yy, xx = dpx[0][k], dpx[1][k]
                print("    ", yy + a, xx, g4[yy, xx], r4[yy, xx])
            print("   got==0 in %d of the differing pixels; want==0 in %d" % (int((g4[dpx][:, :3] == 0).all(axis=1).sum()), int((r4[dpx][:, :3] == 0).all(axis=1).sum())))
        print("rank %d role %d %s: %s" % (rank, role, key, "ok" if bad.size == 0 else "rows differ: %d of %d, first %s last %s" % (bad.size, b - a, bad[:6] + a, bad[-3:] + a)))
