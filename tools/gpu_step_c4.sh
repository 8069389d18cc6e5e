#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -40) > gpurun_out/r2_test4.log; tail -30 gpurun_out/r2_test4.log
timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 256 197 12 0 64 577 16 0 > gpurun_out/r2_attn_bench_tc4.txt 2>&1; cat gpurun_out/r2_attn_bench_tc4.txt
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:attention_tc_bwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_bwd50 python tools/attn_bench.py 1024 50 12 0 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:attention_tc_bwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_bwd77 python tools/attn_bench.py 1024 77 8 1 > /dev/null 2>&1
