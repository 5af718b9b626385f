#!/bin/sh
# First GPU call of the next round: everything that was written without GPU access gets its first
# run and its first numbers, in ONE box session (a call costs minutes of queueing and set-up).
#
#   gpurun --timeout 1500 -- 'sh tools/round2_first_call.sh'          (1 GPU, ~15 min of box time)
#   gpurun --gpus 2 --timeout 900 -- 'sh tools/round2_first_call.sh multi'    (then once with --gpus 8)
#
# Output: gpurun_out/r2_*.txt / .csv; copy what is to be judged into profiles/.
set -x
OUT=gpurun_out
mkdir -p $OUT
if [ "$1" = "multi" ]; then
  N=$(python -c "import torch; print(torch.cuda.device_count())")
  # parity of the opt-in multi-GPU modes first, then all settings timed in one launch
  python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $OUT/r2_multi_gpu_tests_$N.txt 2>&1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
    tools/stage_times_multi_gpu.py > $OUT/r2_stage_times_${N}gpu.txt 2>&1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus $N --steps 30 --warmup 5 > $OUT/r2_bench_${N}gpu.json 2> $OUT/r2_bench_${N}gpu.err
  exit 0
fi
# 1. first GPU run of the tests that sort last (new kernels), then the whole GPU suite
python -m pytest tests/test_zz_split_columns_gpu.py tests/test_zz_phi_converge_gpu.py tests/test_zzz_rows_r16_gpu.py \
  -m gpu -q > $OUT/r2_new_kernel_tests.txt 2>&1
# 2. row-kernel variants, both widths (ids ny+5 .. ny+8)
python tools/rows_variants_ab.py 4096 4096 > $OUT/r2_rows_variants_4096.txt 2>&1
python tools/rows_variants_ab.py 4096 8192 > $OUT/r2_rows_variants_8192.txt 2>&1
# 3. three-phase column stage at 4096 x 4096, two atoms per cell
python tools/stage_times.py 4096 4096 6 > $OUT/r2_stage_times_ndof6.txt 2>&1
# 4. launch list + one full ncu capture of the radix-16 row kernels (never a bench value)
GFMD_B200_ROWS_VARIANT=4104 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $OUT/r2_launches_r16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2_ncu_bench.log 2>&1
GFMD_B200_ROWS_VARIANT=4104 ncu --set full --clock-control none --import-source on -k regex:k_rows_.*_r16 -c 2 \
  -o $OUT/r2_rows_r16 python tools/stage_times.py 4096 4096 > $OUT/r2_ncu_r16.log 2>&1
# 5. the bench as the driver runs it, with the default kernels and with the radix-16 rows
python bench.py > $OUT/r2_bench_default.json 2> $OUT/r2_bench_default.err
GFMD_B200_ROWS_VARIANT=4104 python bench.py --no-cpu-baseline > $OUT/r2_bench_r16.json 2> $OUT/r2_bench_r16.err
# 6. the rest of the GPU suite (long: the compound tests)
python -m pytest tests -m gpu -q -x > $OUT/r2_gpu_suite.txt 2>&1
