#!/bin/bash
# Round-2 evidence for the default bench command on ONE GPU.  Each ncu pass runs only after the plain command exited 0.
# Output: gpurun_out/r2_final/*
O=gpurun_out/r2_final
mkdir -p $O
set -x
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout -s KILL 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err || { tail -5 $O/bench_1gpu.err; exit 1; }
tail -1 $O/bench_1gpu.json | cut -c1-600
timeout -s KILL 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
tail -1 $O/bench_reference.json | cut -c1-400
timeout -s KILL 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-amazon > $O/b_small.log 2>&1 || exit 1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|::k_" -c 600 --csv --log-file $O/launches_reddit_k128.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-amazon > $O/ncu1.log 2>&1
for wl in reddit yelp amazon; do timeout -s KILL 900 bash scripts/r2_traffic.sh $wl 128 | tail -4; done
for kn in k_spmm_rows k_spmm_tc k_spmm_special_cta; do
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -o $O/${kn}_full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-amazon > $O/ncu_$kn.log 2>&1
done
ls -la $O | tail -20
