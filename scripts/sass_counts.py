"""Per-kernel SASS mnemonic counts of the shipped library (evidence for DESIGN.md: TMA bulk copies, FP64 / packed-FP32 /
FP64-tensor instructions):  cuobjdump -sass multi_camera_calibration_b200/libmccba.so | python scripts/sass_counts.py"""
import re
import sys
from collections import Counter, OrderedDict

WATCH = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DMMA", "FFMA2", "FMUL2", "FADD2", "FFMA", "F2F", "MUFU", "SHFL", "WARPSYNC", "LDG", "STG",
         "LDS", "STS", "LDL", "STL", "BAR", "REDUX", "ATOM", "RED"]
kern = OrderedDict()
cur = None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur][op] += 1
        kern[cur]["_total"] += 1
print("%-64s %7s " % ("kernel (demangled prefix)", "instrs") + " ".join("%6s" % w[:6] for w in WATCH))
for k, c in kern.items():
    name = k
    try:
        import subprocess
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        pass
    print("%-64s %7d " % (name[-64:], c["_total"]) + " ".join("%6d" % c[w] for w in WATCH))
