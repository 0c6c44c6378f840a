#!/bin/bash
# usage (GPU box): tools/k2sweep.sh "<bench args>" ...  -> frame ms + per-stage ms for each argument set
for a in "$@"; do
  timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline $a 2> gpurun_out/sweep.err | tail -1 > gpurun_out/sweep.json
  python - "$a" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/sweep.json"))
    print("%-40s frame %.3f ms %7.1f Mray/s e2e %.3f ms | %s" % (sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], " ".join("%s %.3f"%(k[:9],v["ms_per_frame"]) for k,v in d["kernels"].items())))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/sweep.err").read()[-800:])
PY
done
