"""Compact per-kernel table from an `ncu --page raw --csv` export: python tools/ncu_summary.py <raw.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
cols = [("Kernel Name", "kernel", None), ("gpu__time_duration.sum", "us", 1.0), ("launch__grid_size", "grid", 1.0), ("launch__registers_per_thread", "regs", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1.0), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes", 1.0), ("smsp__inst_executed.sum", "Minst", 1e-6),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1lsu%", 1.0), ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1tex%", 1.0),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%", 1.0), ("lts__t_sector_hit_rate.pct", "L2hit%", 1.0), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%", 1.0),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM%", 1.0), ("dram__bytes_read.sum", "rdMB", None), ("dram__bytes_write.sum", "wrMB", None),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long", 1.0), ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst", 1.0)]
def tomb(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
print(" ".join("%-9s" % c[1] for c in cols))
for r in data:
    out = []
    for name, short, scale in cols:
        i = col(name)
        if i is None: out.append("%-9s" % "-"); continue
        if short == "kernel": out.append("%-26s" % r[i].replace("void ", "").split("(")[0][:26]); continue
        if short in ("rdMB", "wrMB"): out.append("%-9.1f" % tomb(r[i], units[i])); continue
        v = float(r[i].replace(",", "")) * scale
        if short == "us" and units[i] == "ms": v *= 1000.0
        if short == "us" and units[i] == "ns": v /= 1000.0
        out.append("%-9.1f" % v)
    print(" ".join(out))
