#!/bin/bash
# TEST INFRASTRUCTURE: builds the reference's own Flex binary (main.cu -> run() -> kernel v36, flex.cu:4010,4761) for sm_100 from
# its unmodified sources where they lie under /root/reference, against stand-ins for the three headers of the LSU "gp" library
# that the reference includes but does not ship (oracle/ref_stubs_gpu/).  Output: oracle/_ref/flex_v36 (git-ignored, travels to
# the GPU box).  The reference's own Makefile is not run (it builds the external gp library first).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "no $REF: keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT/obj_gpu"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-std=c++20 -O3 -w -rdc=true -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ -I$HERE/ref_stubs_gpu -I$REF"
for f in mat main flex DataLoader unitheap tools edgelist adjlist algo_bfs order_deg order_rcm order_gorder; do
  if [ ! -f "$OUT/obj_gpu/$f.o" ]; then
    (cd "$REF" && $NVCC $FLAGS -c "$f.cu" -o "$OUT/obj_gpu/$f.o") &
  fi
done
wait
$NVCC -rdc=true -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ "$OUT"/obj_gpu/*.o -o "$OUT/flex_v36" -lcusparse -lcublas -lpthread
echo "built $OUT/flex_v36"
