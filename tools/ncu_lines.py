#!/usr/bin/env python3
"""Per-source-line executed warp instructions and stall samples of one kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  python tools/ncu_lines.py rep.ncu-rep [--min 0.3]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += int(r[iE]) if r[iE].isdigit() else 0; a[1] += int(r[iS]) if r[iS].isdigit() else 0
tot_e = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total executed warp instructions %d, stall samples %d" % (tot_e, tot_s))
print("%-22s %6s %7s %7s  source" % ("file:line", "", "inst%", "stall%"))
for k in sorted(agg):
    e, s, src = agg[k]
    if 100.0 * e / max(tot_e, 1) >= minpct or 100.0 * s / max(tot_s, 1) >= minpct:
        print("%-22s %6s %6.2f%% %6.2f%%  %s" % ("%s:%d" % k, "", 100.0 * e / tot_e, 100.0 * s / tot_s, src))
