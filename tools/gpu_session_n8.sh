#!/bin/bash
# The 8-GPU gpurun call of round 2 (gpurun --gpus 8): the 1M-vertex row-partitioned solve (BASELINE.json configs[3]) on
# 8 and 4 GPUs with the fused halo / persistent filter kernel.  Short on purpose: the call is charged 8x.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2n8}
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n \
      tools/rowpart_solve.py 316 11 p2p > gpurun_out/${TAG}_rowpart_1m_p2p_n$n.log 2>&1
  echo "rowpart 1M p2p n=$n exit $?"; grep "^{" gpurun_out/${TAG}_rowpart_1m_p2p_n$n.log | tail -1 | cut -c1-600
done
