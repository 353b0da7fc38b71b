timeout -s KILL 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_tcw.py tests/test_gpu_robustness.py -x -q 2>&1 | tail -3
for g in 1 2 4 8; do for cfg in "reddit 128" "reddit 64"; do set -- $cfg; FLEX_HOST_GROUPS=$g timeout -s KILL 600 python bench.py --no-amazon --no-cpu-baseline --workload $1 --k $2 --steps 30 2>gpurun_out/x.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('groups=$g $cfg', 'ms=%.4f e2e_ms=%.3f (device %.3f) e2e=%.0f GF' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['device_ms_per_step'], d['e2e']['value']))" || tail -3 gpurun_out/x.err; done; done
for g in 1 4; do FLEX_HOST_GROUPS=$g timeout -s KILL 600 python bench.py --no-amazon --no-cpu-baseline --workload amazon --k 128 --steps 10 2>gpurun_out/x.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('groups=$g amazon 128', 'ms=%.4f e2e_ms=%.3f (device %.3f) e2e=%.0f GF' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['device_ms_per_step'], d['e2e']['value']))"; done
