#!/bin/bash
# Row-kernel launch configurations (FLEX_ROWS_CFG, fx_spmm.cu:launch_aspt) and L2 eviction hints (FLEX_HINTS) on one workload
run() {
  timeout 600 python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('   GF %.0f  ms %.4f  e2e %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
except Exception as e: print('   failed', e)"
}
WL=${WL:-reddit}; K=${K:-128}
for cfg in ${CFGS:-0 1 2 3}; do
  echo "== $WL k=$K cfg=$cfg hints=1"; FLEX_ROWS_CFG=$cfg FLEX_HINTS=1 run --workload $WL --k $K
done
echo "== $WL k=$K cfg=0 hints=0"; FLEX_ROWS_CFG=0 FLEX_HINTS=0 run --workload $WL --k $K
