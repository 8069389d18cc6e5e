#!/bin/bash
# 8-GPU re-run after the fixes: W = 8 loss parity test, ViT-B-32 bench with the parity block, config 4 (ViT-L-14-336)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 400 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -rs -k "8" 2>&1 | grep -v "UserWarning\|Consider using\|^  assert\|^$" | tail -30) > gpurun_out/r2_multirank_${N}b.log; tail -4 gpurun_out/r2_multirank_${N}b.log
timeout 400 $TR --master-port 29561 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n${N}b.json 2> gpurun_out/r2_bench_n${N}b.err; tail -2 gpurun_out/r2_bench_n${N}b.err
timeout 500 $TR --master-port 29563 bench.py --gpus $N --model ViT-L-14-336 --batch 2048 --grad-checkpointing --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c4_n${N}b.json 2> gpurun_out/r2_c4_n${N}b.err; tail -2 gpurun_out/r2_c4_n${N}b.err | cut -c1-300
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_*_n${N}b.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), round(d['ms_per_step'],1), (d.get('roofline_logits_gemm') or {}).get('frac'), (d.get('parity') or {}), d['roofline_step']['frac'])
PY
