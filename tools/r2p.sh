#!/bin/bash
# round-2 profile captures (one GPU): traffic per stage, full-set summary of one frame, launch list, bench lines
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_traffic.log 2>&1
python tools/ncu_traffic.py gpurun_out/r2p_traffic.csv c3 > gpurun_out/r2p_traffic.json 2> gpurun_out/r2p_traffic.err
# one complete frame, full set (frame 4 of the run: skip 3 warm-up frames x 22 launches)
ncu --set full --clock-control none -s 66 -c 22 -o gpurun_out/r2p_frame python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_frame.log 2>&1
ncu -i gpurun_out/r2p_frame.ncu-rep --page raw --csv > gpurun_out/r2p_frame_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2p_frame_raw.csv > gpurun_out/r2p_frame_summary.txt 2>&1
rm -f gpurun_out/r2p_frame_raw.csv gpurun_out/r2p_frame.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_launches.log 2>&1
python bench.py > gpurun_out/r2p_bench_1gpu_c3.json 2> gpurun_out/r2p_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err
python bench.py --no-cpu-baseline --workload c4 --steps 16 > gpurun_out/r2p_bench_1gpu_c4.json 2>> gpurun_out/r2p_bench.err
python bench.py --no-cpu-baseline --workload c5 --steps 16 > gpurun_out/r2p_bench_1gpu_c5.json 2>> gpurun_out/r2p_bench.err
python bench.py --no-cpu-baseline --orbit 0.5 > gpurun_out/r2p_bench_1gpu_c3_orbit.json 2>> gpurun_out/r2p_bench.err
python bench.py --no-cpu-baseline --frames-in-flight 2 > gpurun_out/r2p_bench_1gpu_c3_fif2.json 2>> gpurun_out/r2p_bench.err
python tools/stage_ms.py gpurun_out/r2p_bench_1gpu_*.json
cat gpurun_out/r2p_frame_summary.txt
tail -3 gpurun_out/r2p_traffic.err
