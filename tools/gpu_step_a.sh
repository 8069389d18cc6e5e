#!/bin/bash
# round-2 GPU check A (1 GPU): loss + GEMM parity tests, loss / GEMM micro-benchmarks, ncu of the peer kernel, short bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_loss.py tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_t3.log; tail -4 gpurun_out/r2_t3.log
timeout 300 python tools/loss_bench.py > gpurun_out/r2_lossbench2.txt 2>&1; tail -16 gpurun_out/r2_lossbench2.txt
timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/r2_gemmbench.txt 2>&1; grep -i "fc  fwd\|proj dgrad\|one block\|all tower" gpurun_out/r2_gemmbench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_peer -s 50 -c 1 -f -o gpurun_out/prof_peer_w8_v3 python tools/loss_bench.py > gpurun_out/ncu_peer2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_peer -s 5 -c 1 -f -o gpurun_out/prof_peer_w1_v3 python tools/loss_bench.py > gpurun_out/ncu_peer3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench3.json')); print(d['value'], d['roofline']['frac'], d['roofline_gemm_family']['frac'], d['roofline_logits_gemm'], d['parity'])"; tail -2 gpurun_out/r2_bench3.err
