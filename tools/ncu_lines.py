"""Join the per-instruction stall samples of one kernel in an ncu report with the line table of the same kernel in a
cubin (nvdisasm -g), and print the share of samples / executed instructions per CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled-name substring> [min share]"""
import collections
import csv
import re
import subprocess
import sys

rep, rx, cubin, mangled = sys.argv[1:5]
min_share = float(sys.argv[5]) if len(sys.argv) > 5 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if r and r[0].startswith("0x"):
        sass.append(r)
base = int(sass[0][0], 16)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0])
tot = ti = 0
for r in sass:
    off = int(r[0], 16) - base
    s, ie = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
    tot += s
    ti += ie
    k = line_of.get(off, ("?", 0))
    agg[k][0] += s
    agg[k][1] += ie
print(f"total samples {tot}, instructions {ti}")
srcs = {}
for (f, ln), (s, ie) in sorted(agg.items(), key=lambda x: x[0]):
    if s / tot >= min_share or ie / ti >= min_share:
        print(f"{f}:{ln:<5d} samples {s / tot:6.1%}  inst {ie / ti:6.1%}")
