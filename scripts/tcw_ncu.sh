#!/bin/bash
# one full ncu capture of the tensor-window kernel (run only after tcw_prof.sh exited 0)
mkdir -p gpurun_out
K=${1:-k_spmm_tc}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/tcw_full -f \
  python bench.py --fmt tcw --tc-threshold ${2:-4} --tc-width ${3:-256} --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/tcw_full.log 2>&1
tail -3 gpurun_out/tcw_full.log
