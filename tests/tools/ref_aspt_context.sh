#!/bin/bash
# Context number: the reference's own ASpT binaries (oracle/_ref/sspmm_{128,32}, built unmodified for sm_100 by
# oracle/ref_build.sh) on the Reddit-shape workload.  Lives under tests/ because it executes something under oracle/.
O=${O:-gpurun_out}
mkdir -p $O
python - <<'PY'
import sys, time; sys.path.insert(0, ".")
from flex_b200 import synth
t=time.time(); rp,c,v=synth.generate("reddit", device="cuda"); synth.write_csv("/tmp/reddit_shape.csv", rp.cpu(), c.cpu(), v.cpu()); print("csv written", time.time()-t)
PY
for bin in sspmm_128:128 sspmm_32:32; do exe=${bin%%:*}; k=${bin##*:}; echo "== $exe reddit-shape k=$k"; timeout 600 oracle/_ref/$exe /tmp/reddit_shape.csv $k 2>&1 | grep -E "GFLOPS|t_pre|errs|vari"; done | tee $O/ref_aspt_reddit.log
