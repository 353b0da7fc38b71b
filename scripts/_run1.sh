timeout -s KILL 900 python -m pytest tests/test_gpu_tcw.py tests/test_gpu_spmm.py -x -q 2>&1 | tail -2
for w in "reddit " "reddit deg" "yelp deg" "amazon "; do set2=($w)
  out=$(ORDER=${set2[1]} STEPS=3 WARM=1 timeout -s KILL 300 python scripts/r2_sweep.py ${set2[0]} 128 4:256:224:1024 2>&1 | tail -1 | sed 's/.*ntc/ntc/')
  echo "${set2[0]} ${set2[1]:-ovo} | $out"
done
