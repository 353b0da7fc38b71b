run() { env "$@" timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$*', round(d['value']), round(d['ms_per_step'],4))"; }
run FLEX_MINB=3
run FLEX_MINB=3 FLEX_NO_TILES=1
run FLEX_MINB=1 FLEX_NO_TILES=1
run FLEX_PANEL_WARPS=32 FLEX_NO_TILES=1
