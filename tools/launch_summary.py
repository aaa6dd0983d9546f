"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total and per-kernel time of the LAST `steps` steps."""
import csv, collections, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr, recs = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    rec = dict(zip(hdr, r))
    try:
        v = float(rec["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = rec["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v * 1e6 if u == "s" else v
    recs.append((rec["Kernel Name"], v))
tail = int(sys.argv[2]) if len(sys.argv) > 2 else len(recs)
recs = recs[-tail:]
d = collections.OrderedDict()
for k, v in recs:
    k = k[:70]
    d.setdefault(k, [0.0, 0])
    d[k][0] += v
    d[k][1] += 1
tot = sum(v[0] for v in d.values())
print(f"{path}: {len(recs)} launches, total {tot:.1f} us")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"  {v[0]:9.1f} us {100*v[0]/tot:5.1f}%  n={v[1]:3d}  avg {v[0]/v[1]:7.1f}  {k}")
