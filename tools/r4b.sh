#!/bin/bash
# usage: tools/r4b.sh <variant> ... — kernel-variant sweep: bench only (image_crc32 must stay 121da8fa), per-stage ms
mkdir -p gpurun_out
for l in "$@"; do
  lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola_$l.so"; [ "$l" = base ] && lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola.so"
  EIDOLA_LIB=$lib timeout 300 python bench.py --steps 32 --warmup 8 --no-cpu-baseline 2> gpurun_out/r4b_$l.err | tail -1 > gpurun_out/r4b_$l.json
  python - "$l" <<'PY'
import json,sys
l=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r4b_%s.json"%l))
    print("%-10s frame %.3f ms %7.1f Mray/s | %s | visits %s | crc %s" % (l, d["ms_per_step"], d["value"], " ".join("%s %.3f"%(k[:7],v["ms_per_frame"]) for k,v in d["kernels"].items()), {k:round(v,2) for k,v in d["visits_per_ray"].items()}, d["image_crc32"]))
except Exception as e:
    print(l, "FAILED", e); print(open("gpurun_out/r4b_%s.err"%l).read()[-600:])
PY
done
