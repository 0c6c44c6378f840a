#!/bin/bash
# usage: r3d.sh <N> "<label>|<bench args>" ...
N=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  L=${spec%%|*}; A=${spec#*|}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --no-cpu-baseline $A > gpurun_out/r3d_n${N}_$L.json 2> gpurun_out/r3d_n${N}_$L.err
  echo "== $L rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|NCCL version" gpurun_out/r3d_n${N}_$L.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3d_n${N}_$L.json"))
    print("$L", "N", d["n_gpus"], round(d["value"]), "Mray/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), "ms; latency", round(d.get("frame_latency_ms") or 0, 3), "crc", d["image_crc32"], (d.get("pipeline") or {}).get("stages"), (d.get("pipeline") or {}).get("peer_MB_per_frame"))
    print("   stage ms per rank", d.get("stage_ms_per_rank"))
except Exception as e:
    print("$L: no line", e)
PY
done
