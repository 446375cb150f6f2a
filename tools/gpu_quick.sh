#!/usr/bin/env bash
# quick GPU call: selected tests (-k expression in $1) and an optional timing command ($2...)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K="$1"; shift
T=${QTAG:-quick}
timeout ${QTIMEOUT:-300} python -m pytest tests -q -x -m gpu -k "$K" > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
if [ $# -gt 0 ]; then timeout 200 "$@" > gpurun_out/${T}_cmd.log 2>&1; tail -12 gpurun_out/${T}_cmd.log; fi
