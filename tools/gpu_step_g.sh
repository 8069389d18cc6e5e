#!/bin/bash
# 8-GPU run: W = 4 / 8 loss parity tests, ViT-B-32 bench (BASELINE config 2) with the parity block, configs 4 and 5
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -rs -k "4 or 8" 2>&1 | grep -v "UserWarning\|Consider using\|^  assert\|^$" | tail -30) > gpurun_out/r2_multirank_$N.log; tail -8 gpurun_out/r2_multirank_$N.log
timeout 600 $TR --master-port 29561 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; grep -c parity gpurun_out/r2_bench_n$N.json; tail -2 gpurun_out/r2_bench_n$N.err
timeout 600 $TR --master-port 29562 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --grad-sync ddp --no-parity > gpurun_out/r2_bench_n${N}_ddp.json 2> gpurun_out/r2_bench_n${N}_ddp.err
timeout 900 $TR --master-port 29563 bench.py --gpus $N --model ViT-L-14-336 --batch 2048 --grad-checkpointing --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c4_n$N.json 2> gpurun_out/r2_c4_n$N.err; tail -2 gpurun_out/r2_c4_n$N.err
timeout 600 $TR --master-port 29564 bench.py --gpus $N --model ViT-B-16 --siglip --batch 1024 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c5_n$N.json 2> gpurun_out/r2_c5_n$N.err; tail -2 gpurun_out/r2_c5_n$N.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_*_n$N*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), round(d['ms_per_step'],1), (d.get('roofline_logits_gemm') or {}).get('frac'), (d.get('parity') or {}).get('ok'), d['roofline_step']['frac'])
PY
