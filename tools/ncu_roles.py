"""Summarise an ncu source-page CSV of conv_tc_kernel by warp role (regions delimited by marker instructions)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
num = lambda x: float(x) if x not in ("", None) else 0.0
def fnum(x):
    try: return float(x)
    except Exception: return 0.0
src = [r[ix['Source']] for r in data]
first = lambda pat: next(i for i, s in enumerate(src) if pat in s)
last = lambda pat: max(i for i, s in enumerate(src) if pat in s)
mma_lo, mma_hi = first('UTCHMMA'), last('UTCBAR')
prod_lo, prod_hi = first('LDGSTS'), last('ARRIVES.LDGSTSBAR')
epi_lo, epi_hi = first('LDTM'), last('ATOMS')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(fnum(r[ix['# Samples']]) for r in data)
print('total samples', tot)
def region(name, lo, hi, pad_lo=60, pad_hi=40):
    lo = max(0, lo - pad_lo); hi = min(len(data), hi + pad_hi)
    s = sum(fnum(r[ix['# Samples']]) for r in data[lo:hi])
    agg = {}
    for r in data[lo:hi]:
        for h in stalls: agg[h] = agg.get(h, 0) + fnum(r[ix[h]])
    top = sorted(agg.items(), key=lambda x: -x[1])[:5]
    print(f"{name}: {s:.0f} samples ({100*s/tot:.1f}%)", [(k[6:], int(v)) for k, v in top])
    for r in sorted(data[lo:hi], key=lambda r: -fnum(r[ix['# Samples']]))[:6]:
        print(f"    {fnum(r[ix['# Samples']]):6.0f} {r[ix['Address']][-5:]} {r[ix['Source']][:70]}")
region('MMA', mma_lo, mma_hi)
region('PRODUCER', prod_lo, prod_hi)
region('EPILOGUE', epi_lo, epi_hi, 80, 120)
