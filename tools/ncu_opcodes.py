"""Warp-stall samples of one kernel of an .ncu-rep, summed by SASS opcode and by stall reason:
python tools/ncu_opcodes.py REPORT KERNEL_REGEX"""
import collections
import csv
import io
import subprocess
import sys

report, kernel = sys.argv[1], sys.argv[2]
text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                      capture_output=True, text=True, check=False).stdout
rows = list(csv.reader(io.StringIO(text)))
start = next(i for i, r in enumerate(rows) if "Source" in r and "Warp Stall Sampling (All Samples)" in r)
hdr = rows[start]
ci, cs = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_op, by_reason, by_op_reason = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
total = 0.0
for r in rows[start + 1:]:
    if r and r[0] == "Kernel Name":  # next captured launch
        break
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[cs] or 0)
    except (ValueError, IndexError):
        continue
    tokens = [t for t in r[ci].split() if not t.startswith("@")]
    op = tokens[0].split(".")[0] if tokens else "?"
    by_op[op] += v
    total += v
    for i, h in reasons:
        try:
            x = float(r[i] or 0)
        except ValueError:
            x = 0.0
        by_reason[h] += x
        by_op_reason[op][h] += x
print(f"total samples {total:.0f}")
for op, v in by_op.most_common(14):
    top = ", ".join(f"{h[6:]} {100 * x / max(v, 1):.0f}%" for h, x in by_op_reason[op].most_common(3))
    print(f"  {op:10s} {100 * v / total:5.1f}%   ({top})")
rt = sum(by_reason.values())
print("by reason: " + ", ".join(f"{h[6:]} {100 * x / rt:.1f}%" for h, x in by_reason.most_common(10)))
