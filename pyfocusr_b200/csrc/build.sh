#!/bin/sh
# Builds libfocusr_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp"
mkdir -p build
pids=""
for f in laplacian spmm sell eigs knn knn_pruned eigsort cpd curvature icp lsap; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libfocusr_b200.so build/laplacian.o build/spmm.o build/sell.o build/eigs.o build/knn.o build/knn_pruned.o build/eigsort.o build/cpd.o build/curvature.o build/icp.o build/lsap.o -lcudart -lgomp
echo "built $(pwd)/libfocusr_b200.so"
