#!/bin/bash
# usage (GPU box): tools/k2ncu.sh [-a "<bench args>"] <variant> ...  -> per-kernel ncu lines (one frame) of the trace stages for each prebuilt libeidola_<variant>.so
ARGS=""
if [ "$1" = "-a" ]; then ARGS="$2"; shift 2; fi
for l in "$@"; do
  lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola_$l.so"; [ "$l" = base ] && lib="$PWD/cis-565-final-vr-raytracer_b200/libeidola.so"
  EIDOLA_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"k_gi_|k_trace_queue|k_indirect|k_direct|k_shadow" -s 60 -c 12 --csv --log-file gpurun_out/k2ncu_$l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/k2ncu_$l.log 2>&1
  python - "$l" <<'PY'
import csv,sys,collections
l=sys.argv[1]
rows=[r for r in csv.reader(open('gpurun_out/k2ncu_%s.csv'%l)) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]),r[ki].split('(')[0][-28:]),{})[r[mi]]=float(r[vi].replace(',',''))
print('==', l)
for (i,k),v in d.items():
    print('%2d %-28s %7.1f us  lanes %4.1f  warps %4.1f%%  issue %4.1f%%  inst %6.1fM  l1wave %4.1f%%  l1hit %4.1f%%' % (i,k,v['gpu__time_duration.sum']/1e3,v['smsp__thread_inst_executed_per_inst_executed.ratio'],v['sm__warps_active.avg.pct_of_peak_sustained_active'],v['smsp__issue_active.avg.pct_of_peak_sustained_active'],v['smsp__inst_executed.sum']/1e6,v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'],v['l1tex__t_sector_hit_rate.pct']))
PY
done
