"""Per-kernel share of one steady-state step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

path, anchor = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
idx = [i for i, n in enumerate(names) if anchor in n]
# a steady-state step: the most common anchor-to-anchor segment length (warm-up builds plans, the tail of a
# bench run times stages separately), taken at its first repetition
lengths = collections.Counter(b - a for a, b in zip(idx, idx[1:]))
step_len = max(lengths, key=lambda n: (lengths[n], n))
first = next(a for a, b in zip(idx[1:], idx[2:]) if b - a == step_len)
seg = range(first, first + step_len)
tot = collections.defaultdict(float)
cnt = collections.Counter()
for i in seg:
    n = re.sub(r"\(.*", "", names[i]).replace("void ", "").replace("rn::", "")
    n = re.sub(r"<.*", "", n) if not any(k in n for k in ("level_kernel", "tile_kernel", "affine", "dense", "pack_")) else n
    tot[n] += vals[i]
    cnt[n] += 1
total = sum(tot.values())
print(f"| kernel | launches | time (us) | share |\n|---|---|---|---|")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| `{n[:70]}` | {cnt[n]} | {v / 1e3:.1f} | {100 * v / total:.1f} % |")
print(f"| **step total** | {sum(cnt.values())} | {total / 1e3:.1f} | 100 % |")
