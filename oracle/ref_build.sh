#!/bin/bash
# Compiles the reference's own host sources, where they lie under /root/reference, into
# oracle/_ref/flexref (CPU-only: g++ against oracle/ref_stubs) and, if nvcc is present, the
# unmodified ASpT binaries oracle/_ref/sspmm_128 / sspmm_32 for sm_100 (run on the GPU box for
# context numbers).  Outputs only under oracle/_ref/ (git-ignored, travels with gpurun).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "no $REF: keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT/obj"
CXX=/usr/bin/g++
[ -x "$CXX" ] || CXX=g++
FLAGS="-std=c++20 -O2 -w -x c++ -I$HERE/ref_stubs -I$REF"
for f in mat DataLoader order_deg order_rcm order_gorder edgelist adjlist algo_bfs unitheap tools; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cu" -nt "$OUT/obj/$f.o" ]; then
    $CXX $FLAGS -c "$REF/$f.cu" -o "$OUT/obj/$f.o" &
  fi
done
wait
$CXX -std=c++20 -O2 -w -I"$HERE/ref_stubs" -I"$REF" "$HERE/ref_driver.cc" "$OUT"/obj/*.o -o "$OUT/flexref"
# the reference's Matrix Market -> CSV converter (data/SuiteSparse/mtx2csr.cc), unmodified
if [ ! -f "$OUT/mtx2csr" ]; then
  $CXX -O2 -w -I"$REF/data/SuiteSparse" "$REF/data/SuiteSparse/mtx2csr.cc" -o "$OUT/mtx2csr" || echo "mtx2csr did not build"
fi
NVCC=/usr/local/cuda/bin/nvcc
if [ -x "$NVCC" ] && [ "${REF_ASPT:-1}" = "1" ]; then
  for b in sspmm_128 sspmm_32; do
    if [ ! -f "$OUT/$b" ]; then
      (cd "$REF/aspt" && $NVCC -std=c++17 -O3 -w -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ $b.cu -o "$OUT/$b") &
    fi
  done
  wait
  # the same sources behind a wrapper that dumps the reference's tile format after it ran (pins A1: ref_aspt_dump.cu)
  for b in sspmm_128 sspmm_32; do
    if [ ! -f "$OUT/${b}_dump" ] || [ "$HERE/ref_aspt_dump.cu" -nt "$OUT/${b}_dump" ]; then
      (cd "$REF/aspt" && $NVCC -std=c++17 -O3 -w -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ -I"$REF/aspt" \
         -DREF_SRC="\"$REF/aspt/$b.cu\"" "$HERE/ref_aspt_dump.cu" -o "$OUT/${b}_dump") &
    fi
  done
  wait
fi
echo "oracle/_ref built: $(ls "$OUT" | tr '\n' ' ')"
