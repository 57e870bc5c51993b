"""Developer helper: aggregate the ncu source page of one kernel into per-opcode executed-instruction counts for
the hot loop (instructions executed more than half as often as the most-executed one) and the top stall sites.
usage: python tools/ncu_hot.py report.ncu-rep <kernel-name-regex> [launch-index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', txt)[1:]
blk = blocks[which]
lines = blk.split("\n")
print("kernel:", lines[0][:100])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
ex = [int(r[ix["Instructions Executed"]]) for r in data]
mx = sorted(ex)[-50]  # the time loop body: (nearly) the largest execution count
hot = [(r, e) for r, e in zip(data, ex) if 0.5 * mx < e]
ops = collections.Counter()
for r, e in hot:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    ops[m.group(1) if m else "?"] += 1
print("hot-loop static instructions:", len(hot), "(loop-body count %d); dynamic instrs per loop trip: %.1f" % (mx, sum(e for _, e in hot) / mx))
print({k: v for k, v in ops.most_common(40)})
samp = sorted(data, key=lambda r: -int(r[ix["Warp Stall Sampling (All Samples)"]]))[:25]
tot = sum(int(r[ix["Warp Stall Sampling (All Samples)"]]) for r in data)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_")]
for r in samp:
    s = int(r[ix["Warp Stall Sampling (All Samples)"]])
    top = sorted(((int(r[ix[c]]), c) for c in stall_cols if r[ix[c]].isdigit()), reverse=True)[:2]
    print(f"{100*s/tot:5.2f}%  {r[ix['Source']].strip()[:70]:70s} {top}")
agg = collections.Counter()
for r in data:
    for c in stall_cols:
        if r[ix[c]].isdigit():
            agg[c] += int(r[ix[c]])
print({k: f"{100*v/sum(agg.values()):.1f}%" for k, v in agg.most_common(10)})
