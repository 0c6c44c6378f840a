#!/bin/bash
# usage (GPU box, N GPUs): tools/mgpu.sh N "<bench args>" ...  -> one line per argument set
N=$1; shift
port=29520
for a in "$@"; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 32 --warmup 8 $a 2> gpurun_out/mgpu.err | tail -1 > gpurun_out/mgpu.json
  cp gpurun_out/mgpu.json "gpurun_out/mgpu_${N}_$(echo $a | tr -c 'a-zA-Z0-9' '_').json"
  python - "$N $a" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/mgpu.json"))
    print("%-44s frame %.3f ms %7.1f Mray/s fps %.0f e2e %.3f ms | %s | exch1 %s" % (sys.argv[1], d["ms_per_step"], d["value"], d["fps"], d["e2e"]["ms_per_step"], " ".join("%s %.3f"%(k[:7],v["ms_per_frame"]) for k,v in d["kernels"].items()), d.get("exchange1_ms")))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/mgpu.err").read()[-600:])
PY
done
