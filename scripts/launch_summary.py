"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per training step."""
import collections, csv, re, sys
path = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]; durs = [float(r['Metric Value']) for r in rows]
starts = [i for i, n in enumerate(names) if 'fuse_feats' in n]
a, b = starts[step], starts[step + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
for r, d in zip(rows[a:b], durs[a:b]):
    n = r['Kernel Name']
    key = re.sub(r'\(.*', '', n)
    key = re.sub(r'^void ', '', key)[:70]
    if 'gemm' in key or 'lstm' in key:
        key += ' ' + r['Grid Size']
    agg[key][0] += 1; agg[key][1] += d
tot = sum(v[1] for v in agg.values())
print(f"step {step}: {b-a} kernels, {tot/1e3:.1f} us total kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]/1e3:9.1f} us {v[0]:4d}x avg {v[1]/v[0]/1e3:7.1f}  {k}")
