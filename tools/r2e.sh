#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2e_pytest.log
for fif in 2 1; do
  timeout 300 python bench.py --steps 32 --warmup 6 --no-cpu-baseline --frames-in-flight $fif > gpurun_out/r2e_bench_fif$fif.json 2> gpurun_out/r2e_bench_fif$fif.err
done
python tools/stage_ms.py gpurun_out/r2e_bench_fif*.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_bench_fif*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['image_crc32'], d['frames_rendered'], d['e2e']['value'], d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
cat gpurun_out/r2e_pytest.log
