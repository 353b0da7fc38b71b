#!/bin/bash
# Results table of the final kernels (tensor-window format, the bench default), one line per named shape / k / ordering
run() {
  timeout 900 python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); tw = d.get('tensor_windows') or {}; r = d['roofline']
    print('   GF %.0f  ms %.4f  frac %.3f  tPre %.2f (x%.1f)  e2e %.0f GF (%.3f ms)  win %.2f ntc %s' % (d['value'], d['ms_per_step'], r['frac'], d.get('tPre_ms') or -1, d.get('tPre_over_tElap') or -1, d['e2e']['value'], d['e2e']['ms_per_step'], (tw.get('win_nnz',0)/max(1,tw.get('win_nnz',0)+tw.get('rest_nnz',1))), tw.get('ntc')))
except Exception as e: print('   failed', e)"
}
for wl in reddit flickr yelp amazon pubmed; do echo "== $wl k=128 tcw"; run --workload $wl --k 128; done
echo "== flickr k=128 tcw order=rcm"; run --workload flickr --k 128 --order rcm
for k in 32 64 256; do echo "== reddit k=$k tcw"; run --workload reddit --k $k; done
echo "== yelp k=32 tcw"; run --workload yelp --k 32
echo "== pubmed k=32 tcw"; run --workload pubmed --k 32
echo "== reddit k=128 aspt"; run --workload reddit --k 128 --fmt aspt
echo "== amazon k=128 aspt"; run --workload amazon --k 128 --fmt aspt
