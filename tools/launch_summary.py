"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list by kernel (optionally only launches whose grid mentions a token)."""
import csv, collections, re, sys
path, token = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).split('::')[-1]
    if token and token not in r[gi]: continue
    tot[name] += float(r[vi].replace(',', '')) / 1e6; cnt[name] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k:20s} {cnt[k]:5d} launches {v:9.3f} ms {100 * v / s:5.1f}%")
print(f"total {s:.3f} ms over {sum(cnt.values())} launches")
