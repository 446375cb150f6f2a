#!/usr/bin/env bash
# consolidated GPU call C: v3 staging + NaN-through-FADD epilogue; parity, smoke, bench (+ reference arm), ncu launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu > gpurun_out/c_fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c_fused_tests.log
tail -4 gpurun_out/c_fused_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/c_bench.json
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/c_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c_all_tests.log
tail -6 gpurun_out/c_all_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
   --log-file gpurun_out/c_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/c_ncu_bench.log 2>&1; echo "ncu rc=$?"
