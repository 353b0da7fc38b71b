timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/r2_build_list.sh r2_b7 reddit 128 | tail -22
