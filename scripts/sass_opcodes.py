#!/usr/bin/env python
"""Per-kernel histogram of the SASS opcodes that matter for the B200 claims (cuobjdump -sass of the built library):
UBLKCP / SYNCS (1-D TMA bulk copies + mbarriers), LDGSTS (cp.async), FFMA2 / FMUL2 / FADD2 (packed FP32), MUFU, MATCH,
REDUX, and the absence of tensor-core opcodes (none of these kernels is a contraction).
Usage: scripts/sass_opcodes.py [libwgsassign_b200.so] > profiles/sass_opcodes_r2.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "wgsassign_b200/libwgsassign_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU.RCP", "MUFU.LG2", "MUFU.EX2", "MUFU", "MATCH", "REDUX",
        "SHFL", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "BAR", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "DFMA", "DADD"]
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
arch = re.search(r"arch = (sm_\w+)", out)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("wgs::", "").replace("void ", "")
        kern = name
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_all"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                hist[kern][w] += 1
                total[w] += 1
                break
print("SASS opcode histogram of %s (%s), %d kernels" % (lib, arch.group(1) if arch else "?", len(hist)))
print("totals:", " ".join("%s=%d" % (w, total[w]) for w in WANT if total[w]))
print("tensor-core opcodes (HMMA / UTC*MMA / LDTM): %d - by design, no kernel of this path is a contraction" % (total["HMMA"] + total["UTCHMMA"] + total["UTCQMMA"] + total["LDTM"]))
print()
cols = ["UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU.RCP", "MUFU.LG2", "MATCH", "REDUX", "SHFL", "LDS", "STS", "BAR"]
print("%-64s %6s " % ("kernel", "insts") + " ".join("%8s" % c for c in cols))
for k, h in hist.items():
    print("%-64s %6d " % (k[:64], h["_all"]) + " ".join("%8d" % h[c] for c in cols))
