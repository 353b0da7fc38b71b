#!/bin/bash
# Measurement campaign over BASELINE.json's configs (1 GPU).  Output: gpurun_out/campaign/*.json
mkdir -p gpurun_out/campaign
O=gpurun_out/campaign
b() { name=$1; shift; timeout 900 python bench.py "$@" 2>$O/$name.err | tail -1 > $O/$name.json; python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "GF=%.0f ms=%.4f frac=%.4f tPre=%.3f e2e=%.0f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["tPre_ms"], d["e2e"]["value"]), "cusparse=%s" % d.get("cusparse_context_gflops"), "cpu=%s" % (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b pubmed_k32 --workload pubmed --k 32 --steps 200 --cusparse
b pubmed_k128 --workload pubmed --k 128 --steps 200 --no-cpu-baseline --cusparse
b flickr_k128 --workload flickr --k 128 --steps 200 --cusparse
b flickr_k128_rcm --workload flickr --k 128 --steps 200 --order rcm --no-cpu-baseline
b flickr_k128_shuffle --workload flickr --k 128 --steps 200 --shuffle --no-cpu-baseline
b flickr_k128_shuffle_rcm --workload flickr --k 128 --steps 200 --shuffle --order rcm --no-cpu-baseline
b reddit_k128 --workload reddit --k 128 --steps 100 --cusparse
b reddit_k32 --workload reddit --k 32 --steps 100 --no-cpu-baseline --cusparse
b reddit_k128_shuffle --workload reddit --k 128 --steps 50 --shuffle --no-cpu-baseline
b reddit_k128_csr --workload reddit --k 128 --steps 50 --fmt csr --no-cpu-baseline
b yelp_k32 --workload yelp --k 32 --steps 100 --cusparse
b yelp_k128 --workload yelp --k 128 --steps 100 --no-cpu-baseline --cusparse
b yelp_k128_deg --workload yelp --k 128 --steps 100 --order deg --no-cpu-baseline
b yelp_k128_gor --workload yelp --k 128 --steps 100 --order gor --no-cpu-baseline
b amazon_k128 --workload amazon --k 128 --steps 20 --no-cpu-baseline
# the reference's own ASpT binary on the headline workload (context number): tests/tools/ref_aspt_context.sh
O=$O bash tests/tools/ref_aspt_context.sh
