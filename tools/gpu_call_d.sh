#!/usr/bin/env bash
# GPU call D: timing experiments on the fused stage (debug switches), ncu --set full of its two kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/prof_fused.py 4096 4096 256 3 0,2,4,6,1,7 > gpurun_out/d_fused_debug_timing.txt 2>&1; echo "timing rc=$?"; cat gpurun_out/d_fused_debug_timing.txt
timeout 120 python tools/prof_fused.py 4096 2048 256 3 0 >> gpurun_out/d_fused_debug_timing.txt 2>&1
timeout 120 python tools/prof_fused.py 2048 4096 256 3 0 >> gpurun_out/d_fused_debug_timing.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sgm_wave -c 2 -o gpurun_out/d_ncu_fused -f \
   python tools/prof_fused.py 1024 4096 256 1 0 > gpurun_out/d_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/ | grep ncu
