"""Per-instruction stall-sample digest of one kernel of an ncu report (needs -lineinfo / --import-source on).
usage: python tools/ncu_stalls.py <report.ncu-rep> <kernel regex> [top-N]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
idx = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hdr_i + 1:]:
    if r and r[0] == "Kernel Name":
        break  # first launch only
    if r and r[0].startswith("0x"):
        sass.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]]) for r in sass)
agg = {h: sum(int(r[idx[h]]) for r in sass) for h in stalls}
print("instructions", len(sass), "samples", tot)
print(", ".join(f"{k[6:]} {v} ({v / tot:.1%})" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
for r in sorted(sass, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
    st = {h[6:]: int(r[idx[h]]) for h in stalls if int(r[idx[h]]) > 0}
    print(r[0][-4:], r[1].strip()[:64].ljust(64), r[idx["# Samples"]].rjust(5),
          sorted(st.items(), key=lambda x: -x[1])[:3])
