#!/bin/bash
# ncu captures of the Flex-format consumers (K2) and of the pillar builder on flickr-shape k=128 -> gpurun_out/r2_k2/
O=gpurun_out/r2_k2
mkdir -p $O
for f in pillar seg tile; do
  kn=k_spmm_panel_acc; [ $f = pillar ] && kn=k_spmm_alpha
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -o $O/${f}_full -f \
    python bench.py --workload flickr --k 128 --fmt $f --steps 3 --warmup 3 --no-cpu-baseline --no-amazon > $O/ncu_$f.log 2>&1
done
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_spmm_alpha|k_spmm_panel_acc|k_pillar|k_pseg|k_diag|k_band|k_seg|k_tile" -c 200 --csv --log-file $O/launches.csv \
  python bench.py --workload flickr --k 128 --fmt pillar --steps 3 --warmup 3 --no-cpu-baseline --no-amazon > $O/ncu_l.log 2>&1
