#!/bin/bash
# Round evidence for the default bench command: bench JSON, ncu launch list, DRAM bytes of the SpMM kernels,
# one full capture of the dominant kernel.  Each ncu pass runs only after the plain command exited 0.
mkdir -p gpurun_out
set -x
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || { tail -5 gpurun_out/bench_default.err; exit 1; }
tail -1 gpurun_out/bench_default.json | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_small.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_default.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_spmm -c 60 --csv \
  --log-file gpurun_out/dram_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmm_rows -s 2 -c 1 -o gpurun_out/rows_full -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tc -s 2 -c 1 -o gpurun_out/tc_full -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmm_special_cta -s 2 -c 1 -o gpurun_out/special_full -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -1 gpurun_out/bench_reference.json | cut -c1-300
