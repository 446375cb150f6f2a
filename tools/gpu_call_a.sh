#!/usr/bin/env bash
# one consolidated GPU call: fused-stage parity, bench (fused / separate), ncu launch list, full GPU suite
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt 2>&1
timeout 200 python tools/fused_diag.py > gpurun_out/a_diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/a_diag.log
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu > gpurun_out/a_fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/a_fused_tests.log
tail -5 gpurun_out/a_fused_tests.log
PB200_FUSE_CENSUS_SGM=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err; echo "bench fused rc=$?"
cat gpurun_out/a_bench_fused.json | cut -c1-600
PB200_FUSE_CENSUS_SGM=1 timeout 420 python -m pytest tests -q -m gpu > gpurun_out/a_all_tests_fused_on.log 2>&1; echo "rc=$?" >> gpurun_out/a_all_tests_fused_on.log
tail -8 gpurun_out/a_all_tests_fused_on.log
PB200_FUSE_CENSUS_SGM=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
   --log-file gpurun_out/a_launches_fused.csv python bench.py --steps 1 --warmup 3 > gpurun_out/a_ncu_bench.log 2>&1; echo "ncu rc=$?"
