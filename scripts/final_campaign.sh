#!/bin/bash
# Round results table: every named shape, ASpT layout vs tensor-window format, plus k sweeps and cuSPARSE context
run() {
  timeout 900 python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); tw = d.get('tensor_windows') or {}
    print('   GF %.0f  ms %.4f  tPre %.2f  e2e %.0f GF (%.3f ms)  win %.2f ntc %s%s' % (d['value'], d['ms_per_step'], d.get('tPre_ms') or -1, d['e2e']['value'], d['e2e']['ms_per_step'], (tw.get('win_nnz',0)/max(1,tw.get('win_nnz',0)+tw.get('rest_nnz',1))), tw.get('ntc'), ('  cusparse %.0f GF' % d['cusparse_gflops']) if 'cusparse_gflops' in d else ''))
except Exception as e: print('   failed', e)"
}
for wl in reddit flickr yelp amazon pubmed; do
  for f in aspt tcw; do echo "== $wl k=128 $f"; run --workload $wl --k 128 --fmt $f; done
done
echo "== reddit k=128 tcw + cusparse context"; run --workload reddit --k 128 --cusparse
for k in 32 64 256; do for f in aspt tcw; do echo "== reddit k=$k $f"; run --workload reddit --k $k --fmt $f; done; done
for o in deg rcm; do echo "== reddit k=128 tcw order=$o"; run --workload reddit --k 128 --order $o; done
