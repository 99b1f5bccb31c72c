"""Prints the instructions with the most warp-stall samples from an `ncu --page source --csv` export."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
    s = int(r[ix['# Samples']] or 0)
    why = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  {r[ix['Source']][:90]:90s} {why}")
