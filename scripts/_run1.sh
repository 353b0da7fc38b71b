for cfg in "512 0 256" "1024 148 256" "1024 148 1024" "768 148 512" "512 296 1024"; do set -- $cfg
  for w in "reddit " "reddit deg" "yelp deg" "amazon "; do set2=($w)
    out=$(FLEX_BUILD_THREADS=$1 FLEX_BUILD_CTAS=$2 FLEX_DETECT_THREADS=$3 ORDER=${set2[1]} STEPS=3 WARM=1 timeout -s KILL 300 python scripts/r2_sweep.py ${set2[0]} 128 4:256:224:1024 2>&1 | tail -1 | sed 's/.*ntc/ntc/')
    echo "threads=$1 ctas=$2 detect=$3 | ${set2[0]} ${set2[1]:-ovo} | $out"
  done
done
