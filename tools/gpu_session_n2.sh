#!/bin/bash
# The 2-GPU gpurun call of round 2 (gpurun --gpus 2): the row-partitioned solve under torchrun (small problem through the
# GPU test, then the 1M-vertex solve in both halo modes) and the bench line at N = 2, which carries rowpart_1m.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2n2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -k "row_partitioned" > gpurun_out/${TAG}_rowpart_test.log 2>&1
echo "rowpart test exit $?"; tail -3 gpurun_out/${TAG}_rowpart_test.log
for mode in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      tools/rowpart_solve.py 316 11 $mode > gpurun_out/${TAG}_rowpart_1m_$mode.log 2>&1
  echo "rowpart 1M $mode exit $?"; grep "^{" gpurun_out/${TAG}_rowpart_1m_$mode.log | tail -1 | cut -c1-700
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
echo "bench n2 exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err
echo "reference arm exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_ref_n2.json
ls -la gpurun_out | tail
