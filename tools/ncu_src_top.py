"""Top stall lines of an `ncu --page source --csv` dump (SASS view):  ncu -i X.ncu-rep --page source --csv > src.csv;
python tools/ncu_src_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
cs, ce, src = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(float(r[cs] or 0) for r in body) or 1.0
print('total samples', tot)
for i in stall_cols:
    s = sum(float(r[i] or 0) for r in body)
    if s > 0.01 * tot:
        print(f'  {hdr[i]:24s} {s / tot:.3f}')
top = sorted(((float(r[cs] or 0), k) for k, r in enumerate(body)), reverse=True)[:n]
for s, k in top:
    r = body[k]
    why = max(stall_cols, key=lambda i: float(r[i] or 0))
    print(f'{s / tot:.3f} #{k:5d} exec={r[ce]:>8s} {hdr[why]:18s} {r[src].strip()[:110]}')
