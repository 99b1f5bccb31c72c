"""Per-source-line warp-stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = ''; out = []; hdr = None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; si = hdr.index('# Samples'); continue
    if hdr and r[0].isdigit():
        stalls = {h: int(v or 0) for h, v in zip(hdr, r) if h.startswith('stall_') and 'Not Issued' not in h and v.isdigit()}
        num = lambda v: int(v) if v.isdigit() else 0
        out.append((num(r[si]), cur_file, r[0], r[1].strip()[:100], sorted(((v, k) for k, v in stalls.items()), reverse=True)[:2]))
tot = sum(o[0] for o in out)
print('total', tot)
for s, f, ln, src, why in sorted(out, reverse=True)[:top]:
    print(f'{s:7d} {100*s/max(tot,1):5.1f}% {f}:{ln:>4s} {src:100s} {why}')
