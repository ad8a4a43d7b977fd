"""Per-kernel SASS opcode evidence from the built library (cuobjdump -sass): the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAPF / UBLKCP = TMA), the legacy tensor path
(HMMA: must be absent), spills (STL / LDL) and the arithmetic mix.  Usage: python tools/sass_histogram.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sglang_fluentllm_b200", "_C", "libb200decode.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCQMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "HMMA", "STL", "LDL",
        "MUFU.EX2", "F2FP", "FFMA2", "FMUL2", "REDUX", "SYNCS", "BAR", "ATOM", "RED"]
cur, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                per[cur][k] += 1
print("library: %s" % os.path.relpath(lib, ROOT))
print("%-96s %6s  %s" % ("kernel (demangled)", "instr", "  ".join(KEYS)))
for name, c in per.items():
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem)
    print("%-96s %6d  %s" % (dem[:96], c["_total"], "  ".join("%*d" % (len(k), c[k]) for k in KEYS)))
