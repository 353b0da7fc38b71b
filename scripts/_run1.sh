bash scripts/r2_k2_profile.sh
mkdir -p gpurun_out/r2_campaign
for w in pubmed flickr; do for f in pillar seg tile; do
timeout -s KILL 600 python bench.py --no-amazon --no-cpu-baseline --workload $w --k 128 --fmt $f --steps 100 2>gpurun_out/x.err | tail -1 > gpurun_out/r2_campaign/${w}_k128_$f.json
python -c "
import json; d=json.load(open('gpurun_out/r2_campaign/${w}_k128_$f.json')); print('$w k=128 $f', 'ms=%.4f GF=%.0f tPre=%.3f ratio=%.1f e2e=%.0f' % (d['ms_per_step'], d['value'], d['tPre_ms'], d['tPre_ms']/d['ms_per_step'], d['e2e']['value']))"
done; done
