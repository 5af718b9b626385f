#!/bin/sh
# Round 2, 1-GPU call: two-CTA-cluster row kernels (ny = 16384), launch list of the 4096^2 step
set -x
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_rows_r16_gpu.py -m gpu -q > $OUT/r2_c4_r16_tests.txt 2>&1
AB_ROUNDS=2 timeout 300 python tools/rows_variants_ab.py 2048 16384 > $OUT/r2_c4_rows_variants_16384.txt 2>&1
GFMD_B200_ROWS_VARIANT=16393 timeout 600 python tools/stage_times.py 16384 16384 3 > $OUT/r2_c4_stage_16384_r16c.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(gather|rows|cols|scatter|finalize|sum)" -c 60 --csv \
  --log-file $OUT/r2_c4_launches_4096.csv python bench.py --grid 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-4096 > $OUT/r2_c4_ncu_bench.log 2>&1
GFMD_B200_ROWS_VARIANT=16393 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rows_.*_r16c -c 2 \
  -o /tmp/r2_c4_rows_r16c python tools/stage_times.py 2048 16384 > $OUT/r2_c4_ncu_r16c.log 2>&1
ncu -i /tmp/r2_c4_rows_r16c.ncu-rep --page raw --csv > $OUT/r2_c4_rows_r16c.raw.csv 2>/dev/null
du -sm $OUT
