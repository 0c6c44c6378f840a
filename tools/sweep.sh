#!/bin/bash
# usage (on the GPU box): tools/sweep.sh "<label>|<nvcc extra flags>" ...   -> per-kernel ms of bench.py for each build variant
for v in "$@"; do
  label="${v%%|*}"; flags="${v#*|}"
  EID_NVCC_EXTRA="$flags" python cis-565-final-vr-raytracer_b200/build.py --force > /dev/null 2> gpurun_out/build_$label.err || { echo "$label BUILD FAILED"; tail -5 gpurun_out/build_$label.err; continue; }
  python bench.py --steps 16 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench_$label.err | tail -1 > gpurun_out/bench_$label.json
  python - "$label" <<'PY'
import json,sys
l=sys.argv[1]
try:
    d=json.load(open("gpurun_out/bench_%s.json"%l))
    print("%-28s frame %.3f ms  %7.1f Mray/s | %s | visits %s" % (l, d["ms_per_step"], d["value"], " ".join("%s %.3f"%(k[:7],v["ms_per_frame"]) for k,v in d["kernels"].items()), {k:round(v,1) for k,v in d["visits_per_ray"].items()}))
except Exception as e:
    print(l, "FAILED", e); print(open("gpurun_out/bench_%s.err"%l).read()[-800:])
PY
done
