#!/usr/bin/env python3
"""Dynamic SASS opcode histogram of one kernel in an .ncu-rep (executed warp instructions per opcode, per tet if --tets).

    python tools/ncu_opcodes.py rep.ncu-rep k_jacobi_tilesN [--tets 10002432]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
tets = int(sys.argv[sys.argv.index("--tets") + 1]) if "--tets" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hist, hdr, name = collections.Counter(), None, None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if name is not None:
            break          # first matching launch only
        name = r[1]
        continue
    if r[0] == "Address":
        hdr = r
        iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
        continue
    if hdr and r[0].startswith("0x"):
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", r[iS])
        if m and r[iE].isdigit():
            op = m.group(1).split(".")
            key = op[0] + ("." + op[1] if len(op) > 1 and op[0] in ("LDS", "STS", "LDG", "STG", "LDGSTS", "MUFU", "BAR", "SHFL") else "")
            hist[key] += int(r[iE])
tot = sum(hist.values())
print("kernel:", name)
print("executed warp instructions: %d%s" % (tot, "  (%.1f thread instructions per tet)" % (tot * 32.0 / tets) if tets else ""))
for op, n in hist.most_common():
    if n * 1000 >= tot:
        print("  %-14s %12d  %5.1f %%%s" % (op, n, 100.0 * n / tot, "  %6.1f / tet" % (n * 32.0 / tets) if tets else ""))
