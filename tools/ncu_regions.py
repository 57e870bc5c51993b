"""Developer helper: where a kernel's time goes along its SASS, in chunks of instructions (stall samples = time of resident warps).
usage: python tools/ncu_regions.py report.ncu-rep <kernel-name-regex> [chunk]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 200
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
blk = re.split(r'(?m)^"Kernel Name",', txt)[1]
lines = blk.split("\n")
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["Warp Stall Sampling (All Samples)"]]) for r in data)
print("instructions", len(data), "samples", tot)
for c0 in range(0, len(data), chunk):
    part = data[c0:c0 + chunk]
    s = sum(int(r[ix["Warp Stall Sampling (All Samples)"]]) for r in part)
    ex = sum(int(r[ix["Instructions Executed"]]) for r in part)
    ops = collections.Counter()
    for r in part:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        ops[m.group(1) if m else "?"] += 1
    print(f"{c0:6d} {100*s/tot:5.1f}%  exec {ex:>12d}  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(5)))
