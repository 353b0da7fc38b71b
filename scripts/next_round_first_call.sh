#!/bin/bash
# First GPU call of the next round (one GPU): everything that was written after this round's GPU budget was spent,
# plus the cheap tPre experiment.  Multi-GPU follow-up (gpurun --gpus 2):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
#     bench.py --gpus 2 --steps 30 --warmup 5     # e2e.path / e2e.replicated_ms show the sharded-input path against fx_spmm_host
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -2
for g in 0 64 96 148 222; do
  echo "== FLEX_BUILD_CTAS=$g (0 = default, 2 per SM)"
  FLEX_BUILD_CTAS=$g SPECS="reddit:128 yelp:128" bash scripts/quick.sh
done 2>&1 | tee gpurun_out/build_ctas.log
