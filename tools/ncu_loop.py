"""Developer helper: time profile of the hot loop of one kernel from an ncu report (best read from a lone-warp capture,
where stall samples are time): per-chunk cycles/instruction, branch list, slowest instructions with their source line.
usage: python tools/ncu_loop.py report.ncu-rep <kernel-regex> <launch-index> <cycles-per-loop-trip>"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, rx, which, cyc = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
blk = re.split(r'(?m)^"Kernel Name",', txt)[1:][which]
lines = blk.split("\n")
print("kernel:", lines[0][:110])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
data = [r for r in rows[1:] if len(r) == len(hdr)]
S = [int(r[ix["Warp Stall Sampling (All Samples)"]]) for r in data]
E = [int(r[ix["Instructions Executed"]]) for r in data]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mx = sorted(E)[-50]  # the loop body: (nearly) the largest execution count
hot = [i for i, e in enumerate(E) if e > 0.5 * mx]
lo, hi = hot[0], hot[-1]
tot = sum(S[i] for i in range(lo, hi + 1))
print(f"loop: instructions {lo}..{hi}, executed per trip {len(hot)}, in range {hi - lo + 1}; samples {tot}")
agg = collections.Counter()
for i in range(lo, hi + 1):
    for c in stall_cols:
        v = data[i][ix[c]]
        if v.isdigit():
            agg[c[6:]] += int(v)
print({k: f"{100 * v / sum(agg.values()):.1f}%" for k, v in agg.most_common(8)})
chunk = 100
for a in range(lo, hi + 1, chunk):
    idx = [i for i in range(a, min(a + chunk, hi + 1))]
    s = sum(S[i] for i in idx)
    ex = sum(1 for i in idx if E[i] > 0.5 * mx)
    n64 = sum(1 for i in idx if E[i] > 0.5 * mx and re.search(r"\b(DFMA|DMUL|DADD|FFMA2?|FMUL2?|FADD2?)\b", data[i][ix["Source"]]))
    top = collections.Counter()
    for i in idx:
        for c in stall_cols:
            v = data[i][ix[c]]
            if v.isdigit():
                top[c[6:]] += int(v)
    t3 = ", ".join(f"{k}:{100 * v / max(s, 1):.0f}" for k, v in top.most_common(3))
    print(f"{a:5d} {s / tot * cyc:7.0f} cyc  executed {ex:3d} (fp {n64:3d})  {s / tot * cyc / max(ex, 1):5.2f} cyc/inst  {t3}")
print("slowest instructions:")
for i in sorted(range(lo, hi + 1), key=lambda i: -S[i])[:40]:
    r = data[i]
    top = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if r[ix[c]].isdigit() and int(r[ix[c]]) > 0), reverse=True)[:2]
    print(f"{i:5d} {S[i] / tot * cyc:6.1f} cyc  {r[ix['Source']].strip()[:70]:70s} {top}")
