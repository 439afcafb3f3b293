"""Summarise an `ncu --page source --print-source sass --csv` export: stall reasons and the hottest SASS instructions.
usage: python profiles/sass_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
data = [r for r in rows[2:] if len(r) >= len(h) - 2 and r[0].strip().isdigit() or (len(r) > 3 and r[0][:2] in ('0x',))]
data = [r for r in rows[2:] if len(r) > h.index('# Samples')]
iS = h.index('# Samples'); iI = h.index('Instructions Executed'); iSrc = h.index('Source')
def I(x):
    try: return int(x)
    except Exception: return 0
tot_s = sum(I(r[iS]) for r in data); tot_i = sum(I(r[iI]) for r in data)
print('total samples', tot_s, 'total warp-inst', tot_i, 'sass lines', len(data))
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
agg = {h[i]: sum(I(r[i]) for r in data if len(r) > i) for i in stall_cols}
print('stalls:', [(k, v) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]])
idx = sorted(range(len(data)), key=lambda i: -I(data[i][iS]))[:top]
for i in sorted(idx):
    r = data[i]
    st = sorted(((h[c][6:], I(r[c])) for c in stall_cols if len(r) > c), key=lambda x: -x[1])[:2]
    print('%5d s=%6d i=%8d  %-72s %s' % (i, I(r[iS]), I(r[iI]), r[iSrc][:72], st))
