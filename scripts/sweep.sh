run() { w=$1; k=$2; extra=$3; shift; shift; shift; env "$@" timeout 200 python bench.py --workload $w --k $k --steps 100 --warmup 5 --no-cpu-baseline $extra 2>/tmp/err.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$w k=$k $extra $*', round(d['value']), round(d['ms_per_step'],4))
except Exception as e:
    print('$w k=$k $extra $* FAILED', e); print(open('/tmp/err.log').read()[-300:])"; }
for w in pubmed:32 pubmed:128 flickr:128 yelp:32 yelp:128; do
  run ${w%%:*} ${w##*:} "" FLEX_PANEL_WARPS=16
  run ${w%%:*} ${w##*:} "" FLEX_PANEL_WARPS=8 FLEX_MINB=6
  run ${w%%:*} ${w##*:} "" FLEX_PANEL_WARPS=8 FLEX_MINB=1
  run ${w%%:*} ${w##*:} "--fmt csr" FLEX_X=1
done
