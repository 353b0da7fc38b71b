#!/bin/bash
run() {
  timeout 600 python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); tw = d.get('tensor_windows') or {}
    print('   GF %.0f  ms %.4f  tPre %.2f  win %.2f ntc %s' % (d['value'], d['ms_per_step'], d.get('tPre_ms') or -1, (tw.get('win_nnz',0)/max(1,tw.get('win_nnz',0)+tw.get('rest_nnz',1))), tw.get('ntc')))
except Exception as e: print('   failed', e)"
}
WL=${WL:-reddit}
echo "== $WL aspt"; run --workload $WL --fmt aspt
for cc in ${CCS:-160 224 320 448}; do for w in ${WS:-256 512}; do
  echo "== $WL tcw chunk_cost=$cc W=$w"; run --workload $WL --fmt tcw --tc-chunk-cost $cc --tc-width $w
done; done
