#!/bin/bash
# round-2 evidence (1 GPU), trimmed: launch list of one bench step + `ncu --set full` captures of the kernel families not yet
# captured from this binary (GEMM / attention single-tile captures: tools/gpu_step_c3.sh, gpu_step_c4.sh, gpu_step_j.sh)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b1024.csv python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2_launches_bench.log 2>&1
timeout 100 $NCU -k regex:gemm_peer -s 50 -c 1 -o gpurun_out/r2f_peer_w8 python tools/loss_bench.py > /dev/null 2>&1
timeout 100 $NCU -k regex:"gemm_tc2_kernel<256, 7>" -s 12 -c 1 -o gpurun_out/r2f_dlogits python tools/loss_bench.py > /dev/null 2>&1
timeout 100 $NCU -k regex:"gemm_tc2_kernel<256, 4>" -s 12 -c 1 -o gpurun_out/r2f_dfeat python tools/loss_bench.py > /dev/null 2>&1
timeout 100 $NCU -k regex:lse_combine -s 5 -c 1 -o gpurun_out/r2f_combine python tools/loss_bench.py > /dev/null 2>&1
timeout 100 $NCU -k regex:attention_tc_bwd_long -s 6 -c 2 -o gpurun_out/r2f_attn_bwd577 python tools/attn_bench.py 64 577 16 0 > /dev/null 2>&1
timeout 100 $NCU -k regex:attention_tc_bwd_kernel -s 3 -c 1 -o gpurun_out/r2f_attn_bwd50 python tools/attn_bench.py 1024 50 12 0 > /dev/null 2>&1
timeout 100 $NCU -k regex:layernorm -s 3 -c 2 -o gpurun_out/r2f_ln python tools/ln_bench.py > /dev/null 2>&1
timeout 150 $NCU -k regex:adamw -s 1 -c 1 -o gpurun_out/r2f_adamw python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
ls -la gpurun_out/r2f_*.ncu-rep | wc -l; tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
