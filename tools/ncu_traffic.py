#!/usr/bin/env python
"""DRAM bytes per launch of the dominant kernels, read from `ncu --set full` reports and written to
profiles/ncu_traffic.json (bench.py's roofline.traffic).  One argument per kernel:  KEY=REPORT:KERNEL_REGEX

    python tools/ncu_traffic.py affine_tma_kernel:c3=gpurun_out/r02_affine_c3.ncu-rep:affine_tma_kernel \\
                                dense_kernel_tp:c2=gpurun_out/r02_dense_c2_v2.ncu-rep:dense_kernel_tp
"""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles", "ncu_traffic.json")
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(report: str, kernel: str) -> int:
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv", "--kernel-name", f"regex:{kernel}"],
                          capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units, first = rows[0], rows[1], rows[2]
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        col = header.index(name)
        total += float(first[col].replace(",", "")) * UNITS[units[col]]
    return int(round(total))


def main() -> None:
    table = {}
    if os.path.exists(OUT):
        with open(OUT, encoding="utf-8") as fh:
            table = json.load(fh)
    table["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch, written by "
                         "tools/ncu_traffic.py from the `ncu --set full` reports named in _sources")
    sources = table.setdefault("_sources", {})
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        report, kernel = rest.rsplit(":", 1)
        table[key] = dram_bytes(report, kernel)
        sources[key] = os.path.relpath(report, REPO)
        print(key, table[key])
    table.pop("dense_kernel:c2", None)
    with open(OUT, "w", encoding="utf-8") as fh:
        json.dump(table, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main()
