#!/bin/bash
# One full ncu capture each of the row kernel and the 512-chunk kernel on the default bench (after the plain command exits 0)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_small.log 2>&1 || { tail -5 gpurun_out/b_small.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmm_rows -s 2 -c 1 -o gpurun_out/rows_full -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmm_special_cta -s 2 -c 1 -o gpurun_out/special_full -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out | tail -8
