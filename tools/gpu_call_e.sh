#!/usr/bin/env bash
# GPU call E: parity + timing after a kernel change (fused tests, SGM parity tests, stage timing, bench)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-e}
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu > gpurun_out/${T}_fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_fused_tests.log
tail -3 gpurun_out/${T}_fused_tests.log
timeout 200 python tools/prof_fused.py 4096 4096 256 3 0,4 > gpurun_out/${T}_timing.txt 2>&1; cat gpurun_out/${T}_timing.txt
timeout 100 python tools/prof_sgm.py 4096 4096 256 3 >> gpurun_out/${T}_timing.txt 2>&1; tail -1 gpurun_out/${T}_timing.txt
timeout 420 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_all_tests.log
tail -4 gpurun_out/${T}_all_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/${T}_bench.json
