"""Summarise an `ncu --page source --csv` dump: instruction mix and the hottest SASS basic blocks (per-pixel counts)."""
import csv, sys
from collections import Counter
src, raw, npix, nshow, nlines = sys.argv[1], sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
rows = list(csv.reader(open(raw)))
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if h in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
             'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
             'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
             'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct'):
        print(h, u, v)
rows = list(csv.reader(open(src)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
tot = sum(int(r[iex] or 0) for r in data)
print('total warp instr', tot, 'per unit', round(tot / npix, 1))
c = Counter()
for r in data:
    op = [o for o in r[isrc].split() if not o.startswith('@')]
    c[op[0].split('.')[0] if op else '?'] += int(r[iex] or 0)
print(' '.join('%s:%.0f' % (k, v / npix) for k, v in c.most_common(22)))
run = []; prev = None; out = []
for r in data:
    ex = int(r[iex] or 0)
    if prev is not None and ex != prev:
        out.append((prev, run)); run = []
    run.append((r[isrc], int(r[isamp] or 0))); prev = ex
out.append((prev, run))
for ex, run in sorted(out, key=lambda t: -t[0] * len(t[1]))[:nshow]:
    print('=== exec/unit %.2f len %d total/unit %.1f samples %d' % (ex / npix, len(run), ex * len(run) / npix, sum(s for _, s in run)))
    for l, s in run[:nlines]:
        print('    %6d %s' % (s, l[:100]))
