#!/bin/bash
# round-2 GPU check B (N GPUs): multi-rank loss parity tests + torchrun bench with the parity block
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
(timeout 1200 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -rs 2>&1 | grep -v "UserWarning\|Consider using\|^  assert\|^$" | tail -40) > gpurun_out/r2_multirank_$N.log; tail -8 gpurun_out/r2_multirank_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 6 --warmup 3 --no-e2e > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$N.json')); print(d['value'], d['ms_per_step'], d['roofline_logits_gemm'], d['parity'])"; tail -3 gpurun_out/r2_bench_n$N.err
