#!/bin/bash
# Build libdgpb.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libdgpb.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 -Xptxas -v"
mkdir -p "$HERE/_obj"
pids=()
for f in common comm dense predict vecchia knn ess ess_small; do
  ( $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" > "$HERE/_obj/$f.log" 2>&1 || { cat "$HERE/_obj/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT" "$HERE"/_obj/{common,comm,dense,predict,vecchia,knn,ess,ess_small}.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
