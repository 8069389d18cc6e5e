#!/bin/bash
# 2-GPU check: attention (compact single-tile backward), p2p / fused-forward-after-idle experiment, N=2 bench with stage times
N=${1:-2}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -30) > gpurun_out/r2_test5.log; tail -5 gpurun_out/r2_test5.log
timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 > gpurun_out/r2_attn_bench_tc5.txt 2>&1; cat gpurun_out/r2_attn_bench_tc5.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tools/p2p_bench.py > gpurun_out/r2_p2p_idle_$N.txt 2>&1; grep "rank 0" gpurun_out/r2_p2p_idle_$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n${N}b.json 2> gpurun_out/r2_bench_n${N}b.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench_n${N}b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline_logits_gemm']); print(d['parity']['ok'])
        for s in d['roofline_gemm_signatures']: print(s)
PY
tail -3 gpurun_out/r2_bench_n${N}b.err
