"""SASS opcode histogram of every kernel of lib/libisac_b200.so (one cuobjdump pass), with the Blackwell-specific mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA, FFMA2/FADD2/FMUL2 = packed FP32x2) called out.
usage: python tools/sass_hist.py [kernel-name regex] > profiles/<file>"""
import collections
import re
import subprocess
import sys

LIB = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200/lib/libisac_b200.so"
rx = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
SPECIAL = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "FFMA2", "FADD2", "FMUL2", "HMMA",
           "DFMA", "DMUL", "DADD", "SYNCS", "LDGSTS", "ACQBULK", "UTCBAR", "UTCCP")
print("# SASS opcode histogram per kernel (static instruction counts of lib/libisac_b200.so, cuobjdump -sass).  'special' lists the")
print("# Blackwell / tensor / TMA / packed-FP32 / FP64 mnemonics present; 'top' the ten most frequent opcodes.")
for fn, h in hist.items():
    name = demangle(fn)
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("isac::", "")
    if rx and not rx.search(name):
        continue
    tot = sum(h.values())
    if tot == 0:
        continue
    sp = ", ".join(f"{k}:{h[k]}" for k in SPECIAL if h.get(k))
    top = ", ".join(f"{k}:{v}" for k, v in h.most_common(10))
    print(f"\n== {name}  ({tot} instructions)\nspecial: {sp or '-'}\ntop: {top}")
