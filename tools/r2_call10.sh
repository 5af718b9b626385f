#!/bin/sh
# Round 2, 1-GPU call: fused atom I/O tests + A/B, bench (both settings), GPU suite
set -x
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_atom_io" > $OUT/r2_c10_fused_io_tests.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline > $OUT/r2_c10_bench_fused.json 2> $OUT/r2_c10_bench_fused.err
BENCH_FUSED_IO=0 timeout 900 python bench.py --no-cpu-baseline > $OUT/r2_c10_bench_unfused.json 2> $OUT/r2_c10_bench_unfused.err
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2_c10_gpu_suite.txt 2>&1
du -sm $OUT
