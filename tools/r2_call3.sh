#!/bin/sh
# Round 2, 1-GPU call: the bench as the driver runs it (both arms), the whole GPU suite, launch list
set -x
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py > $OUT/r2_c3_bench.json 2> $OUT/r2_c3_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/r2_c3_bench_reference.json 2> $OUT/r2_c3_bench_reference.err
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r2_c3_gpu_suite.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/r2_c3_launches.csv \
  python bench.py --grid 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-4096 > $OUT/r2_c3_ncu_bench.log 2>&1
du -sm $OUT
