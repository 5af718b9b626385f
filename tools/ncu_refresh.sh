#!/bin/sh
# Refreshes profiles/r2_ncu_full_summary.json (the DRAM bytes bench.py reports as roofline.traffic):
# one `ncu --set full` launch of each dominant kernel, then tools/ncu_summary.py.  Run on a B200:
#   gpurun --timeout 900 -- 'sh tools/ncu_refresh.sh'
set -x
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rows_.*_r16 -c 2 -o /tmp/ncu_rows_r16 python tools/stage_times.py 4096 4096 > $OUT/ncu_rows_r16.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cols_fused_p2_lr -c 1 -o /tmp/ncu_cols_pipe python tools/stage_times.py 4096 4096 > $OUT/ncu_cols_pipe.log 2>&1
GFMD_B200_ROWS_VARIANT=16393 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rows_.*_r16c -c 2 -o /tmp/ncu_rows_r16c python tools/stage_times.py 2048 16384 > $OUT/ncu_rows_r16c.log 2>&1
for n in rows_r16 cols_pipe rows_r16c; do ncu -i /tmp/ncu_$n.ncu-rep --page raw --csv > $OUT/r2_ncu_$n.raw.csv 2>/dev/null; done
# then, back home:  cp gpurun_out/r2_ncu_*.raw.csv profiles/ && python tools/ncu_summary.py
