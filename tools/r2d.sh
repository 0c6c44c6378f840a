#!/bin/bash
# multi-GPU bench through eid_group: usage tools/r2d.sh <N>
N=${1:-2}
mkdir -p gpurun_out
run() {  # label, extra args
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 4 --no-cpu-baseline $2 > gpurun_out/r2d_n${N}_$1.json 2> gpurun_out/r2d_n${N}_$1.err || { echo "FAILED $1"; tail -20 gpurun_out/r2d_n${N}_$1.err; }
}
run static ""
run orbit_auto "--orbit 0.5 --history auto"
run orbit_always "--orbit 0.5 --history always"
run replicated "--post replicated"
python bench.py --steps 16 --warmup 4 --no-cpu-baseline --orbit 0.5 > gpurun_out/r2d_n1_orbit.json 2> gpurun_out/r2d_n1_orbit.err
python tools/stage_ms.py gpurun_out/r2d_n*.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['image_crc32'], d['frames_rendered'], 'exch1 ms', d.get('exchange1_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
