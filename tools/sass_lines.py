#!/usr/bin/env python3
"""Static SASS budget of one kernel of libtetsim_b200.so, attributed to source lines (needs -lineinfo; no GPU).

    python tools/sass_lines.py k_jacobi_tilesNILi512ELi2ELi2ELi4 [--min 3]

Prints instructions per (file, line) with their opcode mix -- the quick check before spending GPU time on a kernel
change ("did the loop body get shorter?").  Dynamic counts come from tools/ncu_lines.py on an .ncu-rep.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tetsim_b200", "libtetsim_b200.so")
pat = sys.argv[1]
minc = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 1
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, capture_output=True)
    cubins = [f for f in os.listdir(td) if f.endswith(".cubin")]
    dis = ""
    for c in cubins:
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, c)], capture_output=True, text=True).stdout
        if pat in out:
            dis = out
            break
if not dis:
    sys.exit("no kernel matching %r" % pat)
lines = dis.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and pat in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.")), len(lines))
cur, cnt, ops = None, collections.OrderedDict(), collections.defaultdict(collections.Counter)
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        cnt[cur] = cnt.get(cur, 0) + 1
        ops[cur][m.group(2).split(".")[0]] += 1
print(lines[start])
for k in sorted(cnt, key=lambda k: k or ("", 0)):
    if cnt[k] >= minc:
        print("%-26s %4d  %s" % ("%s:%d" % k if k else "?", cnt[k], dict(ops[k])))
print("total", sum(cnt.values()))
