#!/bin/sh
# Round 2, first GPU call (1 GPU): first runs and first numbers of everything written without GPU access.
#   gpurun --timeout 1500 -- 'sh tools/r2_call1.sh'
set -x
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/r2_c1_gpu.txt 2>&1
free -g >> $OUT/r2_c1_gpu.txt; nproc >> $OUT/r2_c1_gpu.txt
# 1. tests that never ran on a GPU
timeout 600 python -m pytest tests/test_split_columns_gpu.py tests/test_phi_converge_gpu.py tests/test_rows_r16_gpu.py \
  -m gpu -q > $OUT/r2_c1_new_kernel_tests.txt 2>&1
# 2. row-kernel variants
timeout 300 python tools/rows_variants_ab.py 4096 4096 > $OUT/r2_c1_rows_variants_4096.txt 2>&1
AB_ROUNDS=1 timeout 300 python tools/rows_variants_ab.py 4096 8192 > $OUT/r2_c1_rows_variants_8192.txt 2>&1
AB_ROUNDS=1 timeout 300 python tools/rows_variants_ab.py 2048 16384 > $OUT/r2_c1_rows_variants_16384.txt 2>&1
# 3. software-pipelined column kernel (second build of the same library)
timeout 200 python tools/stage_times.py 4096 4096 > $OUT/r2_c1_stage_default.txt 2>&1
GFMD_B200_LIB=$PWD/user-gfmd_b200/libgfmd_b200_pipe.so GFMD_B200_COLS_PIPE=1 timeout 200 python tools/stage_times.py 4096 4096 > $OUT/r2_c1_stage_colspipe.txt 2>&1
GFMD_B200_LIB=$PWD/user-gfmd_b200/libgfmd_b200_pipe.so GFMD_B200_COLS_PIPE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/r2_c1_colspipe_tests.txt 2>&1
# 4. two atoms per cell at 4096^2 (three-phase column stage), and the strong-scaling grid on one GPU
timeout 300 python tools/stage_times.py 4096 4096 6 > $OUT/r2_c1_stage_ndof6.txt 2>&1
timeout 600 python tools/stage_times.py 16384 16384 3 > $OUT/r2_c1_stage_16384.txt 2>&1
GFMD_B200_ROWS_VARIANT=16392 timeout 600 python tools/stage_times.py 16384 16384 3 > $OUT/r2_c1_stage_16384_r16.txt 2>&1
# 5. ncu: full capture of the radix-16 row kernels and the pipelined column kernel (never a bench value)
GFMD_B200_ROWS_VARIANT=4104 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rows_.*_r16 -c 2 \
  -o /tmp/r2_c1_rows_r16 python tools/stage_times.py 4096 4096 > $OUT/r2_c1_ncu_r16.log 2>&1
GFMD_B200_LIB=$PWD/user-gfmd_b200/libgfmd_b200_pipe.so GFMD_B200_COLS_PIPE=1 timeout 300 ncu --set full --clock-control none \
  --import-source on -k regex:k_cols_fused_p2_lr -c 1 -o /tmp/r2_c1_cols_pipe python tools/stage_times.py 4096 4096 > $OUT/r2_c1_ncu_pipe.log 2>&1
for n in r2_c1_rows_r16 r2_c1_cols_pipe; do
  ncu -i /tmp/$n.ncu-rep --page raw --csv > $OUT/$n.raw.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/$n.source.csv.gz
done
# 6. the whole GPU suite and the bench as the driver runs them
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r2_c1_gpu_suite.txt 2>&1
timeout 600 python bench.py > $OUT/r2_c1_bench.json 2> $OUT/r2_c1_bench.err
# gpurun_out/ comes back only if it is below 64 MiB
du -sm $OUT; while [ $(du -sm $OUT | cut -f1) -gt 55 ]; do rm -f "$(ls -S $OUT/* | head -1)"; done
