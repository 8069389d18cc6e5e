#!/bin/bash
# round-2 GPU check C (1 GPU): tcgen05 attention forward parity + timings (vs CLIPN_ATTN_TC=0), AdamW parity, full suite
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_optim.py -x -q -m gpu 2>&1 | tail -12) > gpurun_out/r2_attn_test.log; tail -6 gpurun_out/r2_attn_test.log
timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 256 197 12 0 64 577 16 0 > gpurun_out/r2_attn_bench_tc.txt 2>&1; cat gpurun_out/r2_attn_bench_tc.txt
CLIPN_ATTN_TC=0 timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 256 197 12 0 64 577 16 0 > gpurun_out/r2_attn_bench_mma.txt 2>&1; cat gpurun_out/r2_attn_bench_mma.txt
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_full_gpu_tests.log; tail -5 gpurun_out/r2_full_gpu_tests.log
