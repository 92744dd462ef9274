"""Per-opcode and per-region view of one kernel of an .ncu-rep (SASS page): executed warp instructions by mnemonic,
and the hottest contiguous address ranges.   python profiles/sass_profile.py rep.ncu-rep [n_groups]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
groups = float(sys.argv[2]) if len(sys.argv) > 2 else 32768.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ie, ws, src = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
ops, stalls, total, tstall = Counter(), Counter(), 0, 0
body = []
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    n, s = int(r[ie]), int(r[ws])
    text = re.sub(r"^@!?U?P\w+\s+", "", r[src].strip())
    op = text.split()[0].split(".")[0]
    ops[op] += n
    stalls[op] += s
    total += n
    tstall += s
    body.append((n, s, r[src].strip()))
print(f"static instructions {len(body)}, executed {total} ({total / groups:.0f} per group), stall samples {tstall}")
for op, n in ops.most_common(28):
    print(f"  {op:10s} {n:>10d} {100 * n / total:5.1f}%  {n / groups:7.1f}/group   stall {100 * stalls[op] / max(tstall, 1):5.1f}%")
