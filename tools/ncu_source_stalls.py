"""Per-region stall breakdown of an exported ncu source page (ncu -i x.ncu-rep --page source --csv > x.csv).

    python tools/ncu_source_stalls.py x.csv <kernel substring>

Splits the kernel's SASS at USETMAXREG (warp roles) and prints, for the instructions after the register INCREASE (the epilogue
warps), the stall reasons summed over all samples and the 25 hottest instructions."""
import csv
import sys
from collections import Counter


def kernels(path):
    cur, rows, hdr = None, [], None
    for r in csv.reader(open(path)):
        if r and r[0] == 'Kernel Name':
            if cur:
                yield cur, hdr, rows
            cur, rows, hdr = r[1], [], None
        elif r and r[0] == 'Address':
            hdr = r
        elif r and hdr:
            rows.append(r)
    if cur:
        yield cur, hdr, rows


def main():
    path, pat = sys.argv[1], sys.argv[2]
    for name, hdr, rows in kernels(path):
        if pat not in name:
            continue
        col = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        start = next((i for i, r in enumerate(rows) if 'USETMAXREG.TRY_ALLOC' in r[col['Source']]), 0)
        total_all = sum(int(r[col['# Samples']] or 0) for r in rows)
        for title, part in (('low-register roles (producer / MMA issuer / relay)', rows[:start]), ('epilogue warps', rows[start:])):
            tot = Counter()
            for r in part:
                for h in stall_cols:
                    tot[h] += int(r[col[h]] or 0)
            n = sum(int(r[col['# Samples']] or 0) for r in part)
            inst = sum(int(r[col['Instructions Executed']] or 0) for r in part)
            print(f'== {name[:60]} :: {title}: {n} of {total_all} samples, {inst} warp instructions, {len(part)} SASS lines')
            for h, v in tot.most_common(8):
                print(f'   {h:24s} {v:8d}  {100.0 * v / max(n, 1):5.1f} %')
            hot = sorted(part, key=lambda r: -int(r[col['# Samples']] or 0))[:25]
            for r in hot:
                top = max(stall_cols, key=lambda h: int(r[col[h]] or 0))
                print(f'   {int(r[col["# Samples"]] or 0):7d}  {top:18s} {r[col["Source"]].strip()[:90]}')


if __name__ == '__main__':
    main()
