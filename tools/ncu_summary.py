"""Key metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`) as a markdown table."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe %'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'),
        ('launch__registers_per_thread', 'regs'), ('sm__cycles_elapsed.avg.per_second', 'SM clock'), ('launch__grid_size', 'grid'), ('launch__cluster_size', 'cluster')]
cols = [(m, n) for m, n in cols if m in ix]
print('| ' + ' | '.join(n for _, n in cols) + ' |')
print('|' + '---|' * len(cols))
for r in rows[2:]:
    out = []
    for m, _ in cols:
        v = r[ix[m]]
        if m == 'Kernel Name':
            v = v.split('(')[0].replace('void ', '')[:40]
        else:
            try:
                v = f'{float(v.replace(",", "")):.4g} {units[ix[m]]}'.strip()
            except ValueError:
                pass
        out.append(v)
    print('| ' + ' | '.join(out) + ' |')
