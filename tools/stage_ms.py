"""Compact per-stage summary of bench.py JSON lines: python tools/stage_ms.py <file with one JSON line per run> [...]"""
import json, sys
for fn in sys.argv[1:]:
    for line in open(fn):
        line = line.strip()
        if not line.startswith("{"):
            continue
        try:
            d = json.loads(line)
        except Exception:
            continue
        if "kernels" not in d:
            print(fn, {k: d.get(k) for k in ("impl", "value", "ms_per_step", "unavailable")})
            continue
        k = d["kernels"]
        print("%-40s frame %.3f ms  %.0f Mray/s  e2e %.3f ms | K1 %.3f K2 %.3f K3 %.3f K4 %.3f K5 %.3f | n_gpus %d" % (
            fn, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], k["direct_stage"]["ms_per_frame"], k["indirect_stage"]["ms_per_frame"],
            k["denoise_direct"]["ms_per_frame"], k["denoise_indirect"]["ms_per_frame"], k["compose"]["ms_per_frame"], d["n_gpus"]))
