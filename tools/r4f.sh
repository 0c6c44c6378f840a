#!/bin/bash
# the five ray-queue launches of one warm frame, base vs a variant library: duration, instructions, issue, occupancy, lanes
mkdir -p gpurun_out
for l in "$@"; do
  lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola_$l.so"; [ "$l" = base ] && lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola.so"
  EIDOLA_LIB=$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_active.avg \
    --clock-control none --cache-control none -k regex:k_trace_queue --launch-skip 20 -c 5 --csv --log-file gpurun_out/r4f_$l.csv python bench.py --steps 2 --warmup 6 --no-cpu-baseline $BENCH_ARGS > gpurun_out/r4f_$l.log 2>&1
  python - "$l" <<'PY'
import csv, sys, collections
l = sys.argv[1]
rows = list(csv.reader(open("gpurun_out/r4f_%s.csv" % l)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    d.setdefault((r[ii], r[ki][:40]), {})[r[mi]] = r[vi]
for (i, k), m in d.items():
    print(l, i, k, " ".join("%s=%s" % (a.split("__")[1][:22], b) for a, b in m.items()))
PY
done
