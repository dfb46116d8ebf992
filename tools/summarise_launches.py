"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    name = row.get('Kernel Name') or ''
    try: v = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row.get('Metric Unit', '')
    v = v / 1e3 if u in ('nsecond', 'ns') else (v * 1e3 if u in ('msecond', 'ms') else v)
    m = re.search(r'(k_[a-z0-9_]+|gemm_[a-z_]+kernel|[A-Za-z0-9_]+)\s*(<|\()', name)
    key = (m.group(1) if m else name)[:48]
    agg[key][0] += 1; agg[key][1] += v; tot += v
print(f'total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:8.1f} us  {k}')
