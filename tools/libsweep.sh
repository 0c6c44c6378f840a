#!/bin/bash
# usage (on the GPU box): tools/libsweep.sh [-a "<bench args>"] <variant> ...   -> quick parity subset + per-kernel ms of bench.py for each
# prebuilt cis-565-final-vr-raytracer_b200/libeidola_<variant>.so (EID_VARIANT=<variant> python .../build.py); "base" = libeidola.so
ARGS=""
if [ "$1" = "-a" ]; then ARGS="$2"; shift 2; fi
for l in "$@"; do
  lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola_$l.so"; [ "$l" = base ] && lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola.so"
  par=$(EIDOLA_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "traversal or c2 or alpha or small_room or forms" 2>&1 | tail -1)
  EIDOLA_LIB=$lib timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline $ARGS 2> gpurun_out/bench_$l.err | tail -1 > gpurun_out/bench_$l.json
  python - "$l" "$par" <<'PY'
import json,sys
l,par=sys.argv[1],sys.argv[2]
try:
    d=json.load(open("gpurun_out/bench_%s.json"%l))
    print("%-10s frame %.3f ms %7.1f Mray/s | %s | visits %s | parity: %s" % (l, d["ms_per_step"], d["value"], " ".join("%s %.3f"%(k[:7],v["ms_per_frame"]) for k,v in d["kernels"].items()), {k:round(v,1) for k,v in d["visits_per_ray"].items()}, par))
except Exception as e:
    print(l, "FAILED", e, par); print(open("gpurun_out/bench_%s.err"%l).read()[-600:])
PY
done
