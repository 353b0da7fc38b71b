#!/bin/bash
# Quick A/B: one line per (workload, k) with the environment given by the caller
run() {
  timeout 900 python bench.py "$@" --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('   GF %.0f  ms %.4f  tPre %.2f  e2e %.3f ms' % (d['value'], d['ms_per_step'], d.get('tPre_ms') or -1, d['e2e']['ms_per_step']))
except Exception as e: print('   failed', e)"
}
for spec in ${SPECS:-reddit:128 reddit:32 reddit:64 yelp:128 yelp:32 flickr:128 pubmed:32}; do
  wl=${spec%%:*}; k=${spec##*:}
  echo "== $wl k=$k $TAG"; run --workload $wl --k $k --steps ${STEPS:-30}
done
