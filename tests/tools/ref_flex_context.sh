#!/bin/bash
# Context run on the GPU box: the reference's own Flex binary (oracle/_ref/flex_v36, built by oracle/ref_flex_build.sh from the
# unmodified sources + stand-ins for the three unshipped gp headers) next to this repo's consumers of the same formats.
#   tests/tools/ref_flex_context.sh [k]      -> gpurun_out/ref_flex_<input>_k<k>.log, gpurun_out/ref_flex_context.md
k=${1:-128}
mkdir -p gpurun_out/ref_flex
python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
from flex_b200 import synth
rp, c, v = synth.generate("flickr")
synth.write_csv("gpurun_out/ref_flex/flickr.csv", rp, c, v)
PY
for inp in data/pubmed.csv gpurun_out/ref_flex/flickr.csv; do
  name=$(basename $inp .csv)
  (cd gpurun_out/ref_flex && MALLOC_MMAP_THRESHOLD_=4294967296 timeout -s KILL 600 ../../oracle/_ref/flex_v36 ../../$inp $k > ../ref_flex_${name}_k$k.log 2>&1; echo "rc=$?" >> ../ref_flex_${name}_k$k.log)
  echo "== reference flex_v36 $name k=$k"; grep -E "^ROW|rc=|error|Assert|assert" gpurun_out/ref_flex_${name}_k$k.log | cut -c1-400 | head -12
  for f in pillar seg tile aspt tcw; do
    echo "-- flexb200 --format $f"; ./flex_b200/flexb200 $inp $k --format $f --check | grep -E "tElap|GFLOPS|errs"
  done
done
