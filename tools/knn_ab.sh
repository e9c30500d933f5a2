#!/bin/bash
# A/B of the pruned KNN's tile sizes (references per tile x queries per CTA): builds variants of the library that differ
# only in knn_pruned.o (here, before the gpurun call:  bash tools/knn_ab.sh build), then times the bench-shape KNN with
# each on the GPU (bash tools/knn_ab.sh run).  Results are bit-identical by construction (exact search, same tie rule).
set -u
cd "$(dirname "$0")/.."
VARIANTS=${VARIANTS:-"128:128 64:128 64:64 128:64 32:64 256:128"}
CS=pyfocusr_b200/csrc
if [ "${1:-run}" = build ]; then
  mkdir -p $CS/ab
  for v in $VARIANTS; do
    tr=${v%%:*}; tq=${v##*:}
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp \
      -DFB_PK_TR=$tr -DFB_PK_TQ=$tq -c $CS/knn_pruned.cu -o $CS/ab/knn_pruned_${tr}_${tq}.o &&
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $CS/ab/libfocusr_b200_tr${tr}_tq${tq}.so \
      $(ls $CS/build/*.o | grep -v knn_pruned.o) $CS/ab/knn_pruned_${tr}_${tq}.o -lcudart -lgomp && echo "built tr=$tr tq=$tq"
  done
  exit 0
fi
mkdir -p gpurun_out
for v in $VARIANTS; do
  tr=${v%%:*}; tq=${v##*:}
  echo "== references per tile $tr, queries per CTA $tq"
  FOCUSR_B200_LIB=$PWD/$CS/ab/libfocusr_b200_tr${tr}_tq${tq}.so timeout 200 python tools/dense_evidence.py --knn-only --no-micro 2>&1 | tail -1 |
    python -c "
import json,sys
d=json.loads(sys.stdin.read())['knn_pruned_128x15212']
for k,v in d.items(): print('   %s: %.3f ms  %.0f M queries/s  %.0f evaluations/query' % (k, v['ms'], v['queries_per_s']/1e6, v['evaluations_per_query']))"
done 2>&1 | tee gpurun_out/${TAG:-r2}_knn_ab.log
