#!/bin/bash
# A/B of the CTA size of the smoothing pass, compiled for full occupancy (2048 threads per SM: 32 registers per thread);
# VARIANTS = threads per CTA.  (The first A/B of the round varied the resident CTAs of 256 threads: 4 / 6 / 8 -> 20.50 /
# 17.47 / 16.67 ms.)
#   bash tools/smooth_ab.sh build   (here, before the gpurun call)      bash tools/smooth_ab.sh run   (on the GPU)
set -u
cd "$(dirname "$0")/.."
VARIANTS=${VARIANTS:-"128 256 512"}
CS=pyfocusr_b200/csrc
if [ "${1:-run}" = build ]; then
  mkdir -p $CS/ab
  for v in $VARIANTS; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp \
      -DFB_SMOOTH_THREADS=$v -c $CS/sell.cu -o $CS/ab/sell_minb$v.o &&
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $CS/ab/libfocusr_b200_smooth$v.so \
      $(ls $CS/build/*.o | grep -v "/sell.o") $CS/ab/sell_minb$v.o -lcudart -lgomp && echo "built threads=$v"
  done
  exit 0
fi
mkdir -p gpurun_out
for v in $VARIANTS; do
  echo "== smoothing pass with $v threads per CTA"
  FOCUSR_B200_LIB=$PWD/$CS/ab/libfocusr_b200_smooth$v.so timeout 200 python tools/smooth_ab.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/${TAG:-r2}_smooth_ab.log
