#!/bin/bash
# tools/scan_row_variants.sh -- static scan of semantically equivalent forms of the streaming kernel's row (statement / operand order,
# unroll): compiles mm16g_kernel<MODE,K,-6,BORDER> alone for every combination and prints the issue-cost model of its steady loop
# (tools/sass_cost.py), cheapest first.  No GPU needed.    usage: tools/scan_row_variants.sh [K=8] [MODE=2] [BORDER=true]
K=${1:-8}; M=${2:-2}; B=${3:-true}
D=$(mktemp -d); ROOT=$(cd "$(dirname "$0")/.." && pwd)
cat > $D/probe.cu <<EOF
#include "kernels.cuh"
namespace abv { template __global__ void mm16g_kernel<$M, $K, -6, $B>(Mm16Params); }
EOF
one() {
  local f="$1"; local tag=$(echo "$f" | tr -d ' =-' | tr -c 'A-Za-z0-9\n' '_')
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I$ROOT/ambivert_b200/csrc -cubin -o $D/$tag.cubin $D/probe.cu $f 2>/dev/null || { echo "FAIL $f"; return; }
  python $ROOT/tools/sass_cost.py --so $D/$tag.cubin 'mm16g_kernel<' | sed -E "s/^[^{]*//" | python -c "
import sys,ast
d=ast.literal_eval(sys.stdin.read().strip()); print(d['cost_per_cell_pair'], 'instr',d['instr'],'mov',d['mov'],'confl',d['conflicts'],'| $f')"
  rm -f $D/$tag.cubin
}
n=0
for a in 0 1 2; do for o in 0 1 2 3 4 5; do for fe in 0 1; do for u in 6 8 12; do
  one "-DABV_ROW_ADDS=$a -DABV_MAX3_ORDER=$o -DABV_FE_ORDER=$fe -DABV_UNROLL=$u" &
  n=$((n+1)); if [ $((n % 12)) -eq 0 ]; then wait; fi
done; done; done; done > $D/out.txt
wait
sort -n $D/out.txt | head -${4:-15}
rm -rf $D
