"""Developer script: per-kernel time shares from an ncu gpu__time_duration launch list (csv)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    name = r[ki].split('(')[0].replace('void ', '')
    if 'egspr' not in name:
        name = 'torch/other: ' + name[:40]
    v = float(r[vi].replace(',', ''))
    v = v / 1000.0 if r[ui] in ('ns', 'nsecond') else v
    tot[name] += v; cnt[name] += 1
ours = {k: v for k, v in tot.items() if 'egspr' in k}
s = sum(ours.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share of egspr time':>20s}")
for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
    print(f"{k:60s} {cnt[k]:8d} {v:12.1f} {v / cnt[k]:10.1f} {100 * v / s:19.1f}%")
oth = sum(v for k, v in tot.items() if 'egspr' not in k)
print(f"non-egspr kernels (torch fills / copies / flush): {oth:.1f} us in {sum(c for k, c in cnt.items() if 'egspr' not in k)} launches")
