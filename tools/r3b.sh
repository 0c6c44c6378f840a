#!/bin/bash
# multi-GPU bench lines: stage pipeline vs row bands.  usage: r3b.sh <N> [extra bench args]
N=$1; shift
mkdir -p gpurun_out
run() {  # label, args...
  L=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/r3b_n${N}_$L.json 2> gpurun_out/r3b_n${N}_$L.err
  echo "== $L rc=$?"; tail -3 gpurun_out/r3b_n${N}_$L.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3b_n${N}_$L.json"))
    print("$L", "N", d["n_gpus"], round(d["value"]), "Mray/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), "ms; latency", d.get("frame_latency_ms"), "crc", d["image_crc32"], d.get("pipeline"))
    print("   stage ms per rank", d.get("stage_ms_per_rank"))
except Exception as e:
    print("$L: no line", e)
PY
}
run pipeline --mgpu pipeline "$@"
run bands --mgpu bands "$@"
