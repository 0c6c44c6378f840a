"""Per-stage DRAM / L2 traffic of one bench frame from an ncu launch list.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline [--workload c3]
  python tools/ncu_traffic.py gpurun_out/traffic.csv c3      -> profiles/r02_traffic_c3.json

Launches are grouped by stage (kernel name), the STATS instantiations of the one profiling frame are left out, and every sum is
divided by the number of frames in the capture (= k_compose launches).  The file records a hash of the kernel sources so that
bench.py can say whether its `roofline.traffic` still belongs to the kernels it runs."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_of(name):
    n = name.replace("void ", "")
    if "k_direct_stage<1" in n or "k_indirect_stage<1" in n:
        return None                                   # STATS variants (visit counters): the one profiling frame
    if "k_trace_queue<" in n and n.split("k_trace_queue<")[1].split(">")[0].split(",")[1].strip() == "1":
        return None
    if "k_direct_stage" in n or "k_direct_spatial" in n:
        return "direct_stage"
    if "k_gi_" in n or "k_trace_queue" in n or "k_indirect_stage" in n or "k_gi_trace" in n:
        return "indirect_stage"
    if "k_denoise_prep" in n or "k_atrous_tile<0" in n or "k_denoise<0" in n:
        return "denoise_direct"
    if "k_atrous_tile<1" in n or "k_denoise<1" in n:
        return "denoise_indirect"
    if "k_compose" in n:
        return "compose"
    return "other"


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)


def main():
    src, workload = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "c3")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, ui, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        launches.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = (r[vi], r[ui])
    stages = collections.defaultdict(lambda: dict(dram_bytes=0.0, lts_bytes=0.0, ncu_us=0.0, launches=0))
    frames = 0
    for (_, name), m in launches.items():
        st = stage_of(name)
        if "k_compose" in name:
            frames += 1
        if st is None:
            continue
        e = stages[st]
        e["dram_bytes"] += to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
        e["lts_bytes"] += to_bytes(*m["lts__t_bytes.sum"])
        e["ncu_us"] += to_us(*m["gpu__time_duration.sum"])
        e["launches"] += 1
    frames = max(frames, 1)
    # the STATS frame contributes no launches to the trace stages: those stages were summed over frames - 1 frames
    out = {}
    for k, e in stages.items():
        f = frames - 1 if k in ("direct_stage", "indirect_stage") and frames > 1 else frames
        out[k] = {a: (b / f) for a, b in e.items()}
    import bench
    res = {"workload": workload, "frames_in_capture": frames, "kernel_source_hash": bench.kernel_source_hash(),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none, "
                  "bench.py --steps 2 --warmup 3 --no-cpu-baseline; per-frame averages (ncu serialises kernels and flushes caches between "
                  "replays: times are for SHARES only)",
           "stages": out}
    dst = os.path.join(ROOT, "profiles", "r02_traffic_%s.json" % workload)
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
