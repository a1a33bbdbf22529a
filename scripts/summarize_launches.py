"""Per-kernel durations of one chunk from an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u.startswith('n') else (v * 1e3 if u.startswith('m') else v)
    seq.append((r[ki].split('(')[0].replace('<unnamed>::', '')[-44:], v))
idx = [i for i, (n, _) in enumerate(seq) if 'ray_setup' in n]
a, b = idx[0], idx[1]
tot = sum(v for _, v in seq[a:b])
for n, v in seq[a:b]: print(f"{n:46s} {v:9.1f} us {100 * v / tot:5.1f}%")
print(f"chunk total {tot:.1f} us over {b - a} launches")
