#!/bin/bash
# scaling sweep on one 8-GPU box through eid_group
mkdir -p gpurun_out
run() {  # N, label, extra args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 16 --warmup 4 --no-cpu-baseline $3 > gpurun_out/r2g_n$1_$2.json 2> gpurun_out/r2g_n$1_$2.err || { echo "FAILED n$1 $2"; grep -v "^\[W\|^W1" gpurun_out/r2g_n$1_$2.err | tail -15; }
}
run 8 static ""
run 8 orbit_always "--orbit 0.5 --history always"
run 8 replicated "--post replicated"
run 4 static ""
run 2 static ""
run 8 c4 "--workload c4"
run 8 c5 "--workload c5"
python bench.py --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/r2g_n1_static.json 2> gpurun_out/r2g_n1_static.err
python tools/stage_ms.py gpurun_out/r2g_n*.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['image_crc32'], d['frames_rendered'], 'exch1 ms', d.get('exchange1_ms'), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
