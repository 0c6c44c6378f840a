#!/bin/bash
# usage: tools/r4d.sh "<bench args>" ... — one bench.py run per argument string (same library), compact per-stage summary
mkdir -p gpurun_out
i=0
for a in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 32 --warmup 8 --no-cpu-baseline $a 2> gpurun_out/r4d_$i.err | tail -1 > gpurun_out/r4d_$i.json
  python - "$i" "$a" <<'PY'
import json,sys
l,a=sys.argv[1],sys.argv[2]
try:
    d=json.load(open("gpurun_out/r4d_%s.json"%l))
    print("%-44s frame %.3f ms %7.1f Mray/s e2e %.3f | %s | crc %s" % (a, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], " ".join("%s %.3f"%(k[:7],v["ms_per_frame"]) for k,v in d["kernels"].items()), d["image_crc32"]))
except Exception as e:
    print(a, "FAILED", e); print(open("gpurun_out/r4d_%s.err"%l).read()[-600:])
PY
done
