#!/usr/bin/env bash
# Build the UNMODIFIED reference C++ (pybind11 modules) from /root/reference into oracle/_ref/.
# Test infrastructure only: used to pin oracle/ against the reference's own code and as the
# "reference" CPU baseline in bench.py.  Sources are compiled where they lie; nothing is copied.
set -euo pipefail
REF="${PANDORA_REFERENCE:-/root/reference}/src/pandora"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
PY="${PYTHON:-python3}"
if [ ! -d "$REF" ]; then
  echo "reference sources not present at $REF: keeping prebuilt oracle/_ref (if any)"; exit 0
fi
mkdir -p "$OUT"
INC="$($PY -m pybind11 --includes)"
SUF="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
CXXFLAGS="-O3 -std=c++17 -shared -fPIC -Werror"
build() { # name, include dir, sources...
  local name="$1"; local inc="$2"; shift 2
  local target="$OUT/${name}${SUF}"
  local newest; newest="$(ls -t "$@" | head -1)"
  if [ -f "$target" ] && [ "$target" -nt "$newest" ]; then echo "up to date: $target"; return; fi
  echo "g++ $name"
  g++ $CXXFLAGS $INC -I"$inc" "$@" -o "$target"
}
build matching_cost_cpp "$REF/matching_cost/cpp/includes" \
  "$REF/matching_cost/cpp/src/bindings.cpp" "$REF/matching_cost/cpp/src/census.cpp" "$REF/matching_cost/cpp/src/matching_cost.cpp" &
build aggregation_cpp "$REF/aggregation/cpp/includes" \
  "$REF/aggregation/cpp/src/bindings.cpp" "$REF/aggregation/cpp/src/aggregation.cpp" &
# the "next" rows of SURVEY.md 8(f): sub-pixel refinement, cost-volume confidence, criteria
build refinement_cpp "$REF/refinement/cpp/includes" \
  "$REF/refinement/cpp/src/bindings.cpp" "$REF/refinement/cpp/src/refinement.cpp" "$REF/refinement/cpp/src/refinement_tools.cpp" \
  "$REF/refinement/cpp/src/vfit.cpp" "$REF/refinement/cpp/src/quadratic.cpp" &
build cost_volume_confidence_cpp "$REF/cost_volume_confidence/cpp/includes" \
  "$REF/cost_volume_confidence/cpp/src/bindings.cpp" "$REF/cost_volume_confidence/cpp/src/ambiguity.cpp" \
  "$REF/cost_volume_confidence/cpp/src/risk.cpp" "$REF/cost_volume_confidence/cpp/src/interval_bounds.cpp" \
  "$REF/cost_volume_confidence/cpp/src/cost_volume_confidence_tools.cpp" &
build criteria_cpp "$REF/cpp/includes" "$REF/cpp/src/bindings_criteria.cpp" "$REF/cpp/src/criteria.cpp" &
wait
touch "$OUT/__init__.py"
ls -la "$OUT"
