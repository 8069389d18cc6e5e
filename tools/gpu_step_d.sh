#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tools/p2p_bench.py > gpurun_out/r2_p2p_$N.txt 2>&1; grep "rank 0" gpurun_out/r2_p2p_$N.txt; tail -3 gpurun_out/r2_p2p_$N.txt
(timeout 1200 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -rs 2>&1 | grep -v "UserWarning\|Consider using\|^  assert\|^$" | tail -40) > gpurun_out/r2_multirank_$N.log; tail -12 gpurun_out/r2_multirank_$N.log
