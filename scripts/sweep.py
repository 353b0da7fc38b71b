"""Small sweep driver: python scripts/sweep.py  (prints one line per configuration)."""
import json
import os
import subprocess
import sys

CONFIGS = json.loads(os.environ.get("SWEEP", "[]")) or [
    # (workload, k, extra args, env)
    (w, k, extra, env)
    for (w, k) in [("pubmed", 32), ("pubmed", 128), ("flickr", 128), ("yelp", 32), ("yelp", 128)]
    for (extra, env) in [([], {"FLEX_PANEL_WARPS": "16"}), ([], {"FLEX_PANEL_WARPS": "8", "FLEX_MINB": "6"}),
                         ([], {"FLEX_PANEL_WARPS": "8", "FLEX_MINB": "1"}), (["--fmt", "csr"], {})]
]
for w, k, extra, env in CONFIGS:
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "bench.py", "--workload", w, "--k", str(k), "--steps", "100", "--warmup", "5",
                        "--no-cpu-baseline", *extra], capture_output=True, text=True, env=e, timeout=600)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(w, k, extra, env, round(d["value"]), round(d["ms_per_step"], 4), "cusparse", d.get("cusparse_context_gflops"), flush=True)
    except Exception as ex:
        print(w, k, extra, env, "FAILED", r.stderr[-300:], flush=True)
