"""Aggregates an ncu report's source page per CUDA source line: python tools/ncu_lines.py <report.ncu-rep> [top N]
(instructions executed, stall samples, average active threads per instruction)."""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.OrderedDict(); cur = None; tot_i = tot_s = 0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        ii = r.index('Instructions Executed'); si = r.index('# Samples'); ti = r.index('Thread Instructions Executed'); continue
    if r[0] != '':
        try: k = (cur, int(r[0])); inst = int(r[ii]); smp = int(r[si]); th = int(r[ti])
        except Exception: continue
        if inst or smp: agg[k] = (inst, smp, th, r[1].strip()[:100]); tot_i += inst; tot_s += smp
print('total warp-instructions', tot_i, 'samples', tot_s, 'avg threads/inst %.1f' % (sum(v[2] for v in agg.values()) / tot_i))
byf = collections.Counter(); bys = collections.Counter()
for (f, l), (i, s, t, src) in agg.items(): byf[f] += i; bys[f] += s
for f in byf: print('%-28s %5.1f%% inst %5.1f%% samples' % (f, 100 * byf[f] / tot_i, 100 * bys[f] / tot_s))
for (f, l), (i, s, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print('%-12s %4d inst %5.1f%% smp %5.1f%% thr/inst %4.1f | %s' % (f, l, 100 * i / tot_i, 100 * s / tot_s, t / max(i, 1), src))
