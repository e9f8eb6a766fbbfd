"""Top SASS instructions by warp-stall samples from `ncu --page source --csv --print-source sass`."""
import csv
import sys
import collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# multiple kernels concatenated: take the first block
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start = hdr_idx[0]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
hdr = rows[start]
body = rows[start + 1:end]
ci = {k: i for i, k in enumerate(hdr)}
samples = [(int(r[ci['# Samples']] or 0), idx, r) for idx, r in enumerate(body) if len(r) > ci['# Samples']]
total = sum(s for s, _, _ in samples)
print('total samples', total, 'instructions', len(body))
by_op = collections.Counter()
for s, _, r in samples:
    op = r[ci['Source']].split()[0] if r[ci['Source']].split() else '?'
    if op.startswith('@'):
        op = r[ci['Source']].split()[1]
    by_op[op.split('.')[0]] += s
print('by opcode:', [(k, round(100 * v / total, 1)) for k, v in by_op.most_common(14)])
for s, idx, r in sorted(samples, reverse=True)[:top]:
    print(f"{100*s/total:5.1f}%  #{idx:5d}  exec={r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:100]}")
