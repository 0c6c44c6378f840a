#!/bin/bash
# final round-2 captures (one GPU): traffic per stage, full-set summary of one frame, cold + warm launch lists, smoke, bench lines
mkdir -p gpurun_out
P=gpurun_out/r5
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv --log-file ${P}_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_traffic.log 2>&1
python tools/ncu_traffic.py ${P}_traffic.csv c3 > ${P}_traffic.json 2> ${P}_traffic.err
ncu --set full --clock-control none -s 60 -c 22 -o ${P}_frame python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_frame.log 2>&1
ncu -i ${P}_frame.ncu-rep --page raw --csv > ${P}_frame_raw.csv 2>/dev/null
python tools/ncu_summary.py ${P}_frame_raw.csv > ${P}_frame_summary.txt 2>&1
rm -f ${P}_frame_raw.csv ${P}_frame.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file ${P}_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_launches_warm.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -2 ${P}_smoke.log | cut -c1-300
python bench.py > ${P}_bench_1gpu_c3.json 2> ${P}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
python bench.py --no-cpu-baseline --workload c4 --steps 16 > ${P}_bench_1gpu_c4.json 2>> ${P}_bench.err
python bench.py --no-cpu-baseline --workload c5 --steps 16 > ${P}_bench_1gpu_c5.json 2>> ${P}_bench.err
python bench.py --no-cpu-baseline --orbit 0.5 > ${P}_bench_1gpu_c3_orbit.json 2>> ${P}_bench.err
python tools/stage_ms.py ${P}_bench_1gpu_*.json ${P}_bench_reference.json
cat ${P}_frame_summary.txt | cut -c1-250
tail -3 ${P}_traffic.err
