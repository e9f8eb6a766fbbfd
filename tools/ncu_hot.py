"""Hottest SASS lines of one kernel from an .ncu-rep:  python tools/ncu_hot.py REPORT KERNEL_REGEX [TOP]"""
import csv
import io
import subprocess
import sys

report, kernel = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                      capture_output=True, text=True, check=False).stdout
rows = list(csv.reader(io.StringIO(text)))
blocks, current = [], None
for row in rows:
    if row and row[0] == "Kernel Name":
        current = {"name": row[1], "rows": []}
        blocks.append(current)
    elif current is not None:
        current["rows"].append(row)
for block in blocks[:1]:
    hdr = block["rows"][0]
    ci, cs, cx = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = []
    for row in block["rows"][1:]:
        try:
            data.append((float(row[cs] or 0), row[cx], row[ci]))
        except (ValueError, IndexError):
            continue
    total = sum(d[0] for d in data) or 1.0
    print(block["name"], "samples", total, "sass lines", len(data))
    for samples, executed, source in sorted(data, key=lambda d: -d[0])[:top_n]:
        print(f"{samples / total * 100:5.1f}%  x{executed:>8}  {source[:110]}")
