"""SASS opcode histogram of every kernel in the shipped library (cuobjdump -sass): the opcodes that prove the
FP64 tensor path (DMMA), bulk TMA copies (UBLKCP), cp.async (LDGSTS), mbarriers (SYNCS) — and that no
tcgen05 / TMEM opcode is expected for an FP64 workload.  python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, "ramannoodle_b200", "libramannoodle_b200.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "LDGSTS", "SYNCS", "UTMALDG", "UTCHMMA", "LDTM", "LDS", "STS", "LDG", "STG", "RED",
        "ATOMG", "MUFU", "FRND", "BAR"]
kernels, name = collections.OrderedDict(), None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        kernels[name][m.group(1)] += 1
demangled = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True, check=False).stdout.splitlines()
if len(demangled) != len(kernels):
    demangled = list(kernels)
print(f"# {os.path.relpath(lib, REPO)}: {len(kernels)} kernels (sm_100a); opcode counts per kernel\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
totals = collections.Counter()
for (mangled, counts), pretty in zip(kernels.items(), demangled):
    cut = pretty.rfind(">(")
    short = (pretty[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", pretty)).replace("void ", "")
    short = short.replace("(int)", "").replace("(bool)", "")
    short = short if len(short) < 110 else short[:107] + "..."
    print(f"| `{short}` | {sum(counts.values())} | " + " | ".join(str(counts.get(k, 0)) for k in KEYS) + " |")
    totals.update(counts)
print(f"| **all kernels** | {sum(totals.values())} | " + " | ".join(str(totals.get(k, 0)) for k in KEYS) + " |")
absent = [k for k in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "HMMA", "IMMA") if totals.get(k, 0) == 0]
print(f"\nabsent (as expected: tcgen05 / TMEM have no f64 kind; TMA is used in its bulk, non-tensor form): {', '.join(absent)}")
