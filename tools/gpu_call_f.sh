#!/usr/bin/env bash
# GPU call F: register CBCA kernel -- parity tests, timing vs the staged kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin_api.py -q -m gpu -k "cbca or disparity_host or host_entry or banded or pipeline" > gpurun_out/f_cbca_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_cbca_tests.log
tail -12 gpurun_out/f_cbca_tests.log
timeout 200 python tools/prof_cbca.py 2048 2048 192 5 > gpurun_out/f_cbca_timing.txt 2>&1; cat gpurun_out/f_cbca_timing.txt
timeout 100 python tools/prof_cbca.py 1024 1024 128 5 >> gpurun_out/f_cbca_timing.txt 2>&1; tail -3 gpurun_out/f_cbca_timing.txt
