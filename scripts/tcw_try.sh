#!/bin/bash
# TCW format: parity tests, then Reddit-shape timing against the ASpT path for a few plan parameters.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcw.py -x -q -m gpu 2>&1 | tail -15
for cfg in "aspt 0 0" "tcw 4 512" "tcw 4 256" "tcw 8 512" "tcw 3 1024" "tcw 6 1024"; do
  set -- $cfg
  echo "== $cfg"
  timeout 600 python bench.py --fmt $1 --tc-threshold $2 --tc-width $3 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l); continue
    print({k: d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d.get('config',{}).get('tPre_ms'), d.get('check'))
"
done
