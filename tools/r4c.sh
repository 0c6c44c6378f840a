#!/bin/bash
# L2 persistence of the acceleration structure: EIDOLA_L2_PERSIST = 0 (off) / 1 (nodes) / 2 (triangles) / 3 (span of both)
mkdir -p gpurun_out
for m in 0 1 2 3 0; do
  EIDOLA_L2_PERSIST=$m timeout 300 python bench.py --steps 32 --warmup 8 --no-cpu-baseline 2> gpurun_out/r4c_$m.err | tail -1 > gpurun_out/r4c_$m.json
  python - "$m" <<'PY'
import json,sys
l=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r4c_%s.json"%l))
    print("persist=%-3s frame %.3f ms %7.1f Mray/s e2e %.3f | %s | crc %s" % (l, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], " ".join("%s %.3f"%(k[:7],v["ms_per_frame"]) for k,v in d["kernels"].items()), d["image_crc32"]))
except Exception as e:
    print(l, "FAILED", e); print(open("gpurun_out/r4c_%s.err"%l).read()[-600:])
PY
done
