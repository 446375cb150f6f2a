#!/usr/bin/env bash
# One consolidated GPU call (run under gpurun): full GPU test suite, smoke, bench (+ reference arm), stage bench and the
# ncu launch list of the bench.  Outputs under gpurun_out/<tag>_*.
set -u
T=${1:-check}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_all_tests.log
tail -4 gpurun_out/${T}_all_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/${T}_bench.json
if [ "${2:-}" = "full" ]; then
timeout 300 python tools/stage_bench.py --json gpurun_out/${T}_stage_bench.json > gpurun_out/${T}_stage_bench.log 2>&1; echo "stage bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file gpurun_out/${T}_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
