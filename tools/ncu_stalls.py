"""Developer script: summarise the per-instruction stall samples of one kernel of an .ncu-rep
(ncu --page source --csv), grouped by SASS opcode and by stall reason.
    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel-index>"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, kid = sys.argv[1], int(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name",')[1:]
blk = blocks[kid]
lines = blk.split("\n")
print("kernel:", lines[0][:140])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
isamp = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_op, by_reason = Counter(), Counter()
total = 0
top = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    s = int(r[isamp] or 0)
    total += s
    op = r[1].split()[0] if r[1].split() else "?"
    if op.startswith("@"):
        op = r[1].split()[1]
    by_op[op.split(".")[0]] += s
    for i in stall_cols:
        by_reason[hdr[i]] += int(r[i] or 0)
    top.append((s, r[1].strip()[:90], {hdr[i]: int(r[i]) for i in stall_cols if int(r[i] or 0) > 0}))
print("total samples", total)
print("by opcode:", [(k, v, "%.1f%%" % (100 * v / total)) for k, v in by_op.most_common(12)])
print("by reason:", [(k, v, "%.1f%%" % (100 * v / total)) for k, v in by_reason.most_common(10)])
top.sort(key=lambda t: -t[0])
for s, src, d in top[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    print("%6d %5.1f%%  %-90s %s" % (s, 100 * s / total, src, sorted(d.items(), key=lambda kv: -kv[1])[:3]))
