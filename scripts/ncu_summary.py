"""One line per `ncu --set full --csv --page raw` capture: duration, DRAM traffic and bandwidth, occupancy, issue slots."""
import csv, sys, json
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum']
SCALE = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
out = {}
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    if len(rows) < 3:
        print(path, 'no kernel captured'); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            try:
                d[h] = float(v.replace(',', '')) * SCALE.get(u, 1.0)
            except ValueError:
                pass
        if h == 'Kernel Name':
            d['kernel'] = v.split('(')[0]
    t = d['gpu__time_duration.sum']
    by = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    rec = dict(kernel=d.get('kernel'), us=round(t * 1e6, 1), dram_MB=round(by / 1e6, 1), dram_GBps=round(by / t / 1e9, 0),
               regs=int(d['launch__registers_per_thread']), warps_active_pct=round(d['sm__warps_active.avg.pct_of_peak_sustained_active'], 1),
               issue_active_pct=round(d['smsp__issue_active.avg.pct_of_peak_sustained_active'], 1),
               l2_hit_pct=round(d.get('lts__t_sector_hit_rate.pct', 0), 1), l1_hit_pct=round(d.get('l1tex__t_sector_hit_rate.pct', 0), 1),
               fp64_pipe_pct=round(d.get('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 0), 1),
               grid=int(d['launch__grid_size']), block=int(d['launch__block_size']), warp_instr=int(d['smsp__inst_executed.sum']))
    out[path] = rec
    print(json.dumps(rec))
