#!/bin/bash
# Round-2 measurement campaign over BASELINE.json's configs and the orderings feeding the tensor windows (1 GPU).
# Output: gpurun_out/r2_campaign/*.json + one summary line each.
mkdir -p gpurun_out/r2_campaign
O=gpurun_out/r2_campaign
b() { name=$1; shift; timeout -s KILL 900 python bench.py --no-amazon "$@" 2>$O/$name.err | tail -1 > $O/$name.json; python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    tw=d.get("tensor_windows") or {}
    print("%-26s GF=%8.0f ms=%.4f frac=%.4f tPre=%7.3f tPre/tElap=%5.1f e2e=%7.0f win=%s ktimes=%s cpu=%s" % (sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["tPre_ms"],
          d["tPre_ms"]/d["ms_per_step"], d["e2e"]["value"], ("%.2f" % (tw["win_nnz"] / max(1, tw["win_nnz"] + tw["rest_nnz"]))) if tw else "-", d["roofline"].get("kernel_ms"), (d.get("cpu_baseline") or {}).get("value")), flush=True)
except Exception as e:
    print(sys.argv[2], "FAILED", e, flush=True)
PY
}
b pubmed_k32 --workload pubmed --k 32 --steps 200
b pubmed_k128 --workload pubmed --k 128 --steps 200 --no-cpu-baseline
b flickr_k128 --workload flickr --k 128 --steps 200
b flickr_k128_rcm --workload flickr --k 128 --steps 200 --order rcm --no-cpu-baseline
b reddit_k32 --workload reddit --k 32 --steps 100 --no-cpu-baseline
b reddit_k64 --workload reddit --k 64 --steps 100 --no-cpu-baseline
b reddit_k128 --workload reddit --k 128 --steps 100
b reddit_k128_aspt --workload reddit --k 128 --steps 100 --fmt aspt --no-cpu-baseline
b reddit_k128_csr --workload reddit --k 128 --steps 50 --fmt csr --no-cpu-baseline
b reddit_k128_deg --workload reddit --k 128 --steps 50 --order deg --no-cpu-baseline
b reddit_k128_rcm --workload reddit --k 128 --steps 50 --order rcm --no-cpu-baseline
b reddit_k128_rbt --workload reddit --k 128 --steps 50 --order rbt --no-cpu-baseline
b reddit_k128_shuffle --workload reddit --k 128 --steps 50 --shuffle --no-cpu-baseline
b reddit_k128_shuffle_rbt --workload reddit --k 128 --steps 50 --shuffle --order rbt --no-cpu-baseline
b yelp_k32 --workload yelp --k 32 --steps 100 --no-cpu-baseline
b yelp_k128 --workload yelp --k 128 --steps 100 --no-cpu-baseline
b yelp_k128_deg --workload yelp --k 128 --steps 100 --order deg --no-cpu-baseline
b yelp_k128_gor --workload yelp --k 128 --steps 100 --order gor --no-cpu-baseline
b amazon_k128 --workload amazon --k 128 --steps 20 --no-cpu-baseline
b flickr_k128_pillar --workload flickr --k 128 --steps 50 --fmt pillar --no-cpu-baseline
b flickr_k128_seg --workload flickr --k 128 --steps 50 --fmt seg --no-cpu-baseline
b flickr_k128_tile --workload flickr --k 128 --steps 50 --fmt tile --no-cpu-baseline
b pubmed_k128_pillar --workload pubmed --k 128 --steps 100 --fmt pillar --no-cpu-baseline
b pubmed_k128_seg --workload pubmed --k 128 --steps 100 --fmt seg --no-cpu-baseline
b pubmed_k128_tile --workload pubmed --k 128 --steps 100 --fmt tile --no-cpu-baseline
