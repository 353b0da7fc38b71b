#!/bin/bash
# ASpT vs tensor-window format across the named shapes (1 GPU)
run() {
  timeout 900 python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('   GF %.0f  ms %.4f  tPre %.2f' % (d['value'], d['ms_per_step'], d.get('tPre_ms') or -1))
except Exception as e: print('   failed', e)"
}
for wl in ${WORKLOADS:-reddit flickr yelp amazon}; do
  for k in ${KS:-128}; do
    echo "== $wl k=$k aspt"; run --workload $wl --k $k --fmt aspt
    for cfg in ${CFGS:-0:0:0}; do
      IFS=: read -r a b c <<< "$cfg"; set -- $a $b $c
      echo "== $wl k=$k tcw T=$1 W=$2 gain=$3"; run --workload $wl --k $k --fmt tcw --tc-threshold $1 --tc-width $2 --tc-min-gain $3
    done
  done
done
