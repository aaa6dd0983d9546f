"""Summarise an ncu source-page CSV of conv_tc2_kernel: warp-state samples by warp role + hottest instructions.
    ncu -i X.ncu-rep --page source --csv > /tmp/src.csv ; python tools/ncu_roles2.py /tmp/src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(x):
    try: return float(x)
    except Exception: return 0.0
src = [r[ix['Source']] for r in data]
first = lambda pat: next(i for i, s in enumerate(src) if pat in s)
last = lambda pat: max(i for i, s in enumerate(src) if pat in s)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(f(r[ix['# Samples']]) for r in data)
print('total samples', tot)
mma_lo, mma_hi = first('UTCHMMA') - 140, last('UTCBAR') + 30
prod_hi = last('UTMALDG') + 25
epi_lo = first('LDTM') - 60
tail_lo = last('UCGABAR_ARV') - 12
regions = [('setup', 0, mma_lo), ('MMA', mma_lo, mma_hi), ('PRODUCER', mma_hi, prod_hi), ('EPILOGUE', prod_hi, tail_lo), ('teardown', tail_lo, len(data))]
for name, lo, hi in regions:
    s = sum(f(r[ix['# Samples']]) for r in data[lo:hi])
    agg = {}
    for r in data[lo:hi]:
        for h in stalls: agg[h] = agg.get(h, 0) + f(r[ix[h]])
    top = sorted(agg.items(), key=lambda x: -x[1])[:5]
    print(f"{name}: {s:.0f} samples ({100*s/tot:.1f}%)", [(k[6:], int(v)) for k, v in top])
    for i in sorted(range(lo, hi), key=lambda i: -f(data[i][ix['# Samples']]))[:7]:
        r = data[i]
        print(f"    {f(r[ix['# Samples']]):6.0f} [{i}] {r[ix['Source']][:70]}")
