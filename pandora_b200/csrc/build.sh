#!/usr/bin/env bash
# Build libpandora_b200.so (sm_100a) in-tree: pandora_b200/_lib/libpandora_b200.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../_lib"
mkdir -p "$OUT" "$HERE/_obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden -I"$HERE/../../include" -I"$HERE")
[ -n "${PB200_PTXAS_V:-}" ] && FLAGS+=(-Xptxas -v)
pids=()
for f in census wta sad_zncc cbca sgm sgm_narrow sgm_wave1 masks validation refinement confidence api; do
  src="$HERE/$f.cu"; obj="$HERE/_obj/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/common.cuh" -nt "$obj" ] || [ "$HERE/sgm_common.cuh" -nt "$obj" ] || [ "$HERE/sgm_packed.cuh" -nt "$obj" ] || [ "$HERE/../../include/pandora_b200.h" -nt "$obj" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libpandora_b200.so" "$HERE"/_obj/{census,wta,sad_zncc,cbca,sgm,sgm_narrow,sgm_wave1,masks,validation,refinement,confidence,api}.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT/libpandora_b200.so"
