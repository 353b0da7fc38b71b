timeout -s KILL 900 python -m pytest tests/test_gpu_tcw.py -x -q 2>&1 | tail -5
bash scripts/r2_build_list.sh r2_b6 reddit 128 | tail -22
for cfg in "reddit 32 tcw" "amazon 128 tcw"; do set -- $cfg; timeout -s KILL 600 python bench.py --no-amazon --no-cpu-baseline --workload $1 --k $2 --fmt $3 --steps 30 2>gpurun_out/x.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', 'ms=%.4f tPre=%.3f ratio=%.1f e2e=%.0f GF=%.0f' % (d['ms_per_step'], d['tPre_ms'], d['tPre_ms']/d['ms_per_step'], d['e2e']['value'], d['value']))" || tail -3 gpurun_out/x.err; done
