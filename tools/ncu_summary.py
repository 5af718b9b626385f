"""profiles/r2_ncu_*.raw.csv (ncu --page raw --csv) -> profiles/r2_ncu_full_summary.json: duration, DRAM bytes,
pipe utilisation, stall reasons per captured kernel.  bench.py reads the DRAM bytes as roofline.traffic."""
import csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
FILES = (('profiles/r2_ncu_rows_r16.raw.csv', '4096x4096 ndof 3'), ('profiles/r2_ncu_cols_pipe.raw.csv', '4096x4096 ndof 3'),
         ('profiles/r2_ncu_rows_r16c.raw.csv', '2048x16384 ndof 3'))
out = {}
for fn, grid in FILES:
    rows = list(csv.reader(open(os.path.join(ROOT, fn))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
        d = {'grid': grid, 'capture': 'ncu --set full --clock-control none (one launch, cold caches)', 'source': fn}
        for w in WANT:
            if w in hdr:
                d[w] = '%s %s' % (r[hdr.index(w)], units[hdr.index(w)])
        out[name] = d
json.dump(out, open(os.path.join(ROOT, 'profiles', 'r2_ncu_full_summary.json'), 'w'), indent=1)
print('\n'.join('%s: %s, read %s, written %s' % (k, v.get('gpu__time_duration.sum'), v.get('dram__bytes_read.sum'),
                                                 v.get('dram__bytes_write.sum')) for k, v in out.items()))
