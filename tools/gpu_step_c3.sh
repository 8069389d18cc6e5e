#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_optim.py tests/test_gpu_loss.py tests/test_gpu_gemm.py "tests/test_gpu_model.py::test_vitl14_336_full_depth_vs_reference_fixture" tests/test_gpu_model.py::test_accum_freq_feature_cache_algorithm_on_native_objects -q -m gpu -s --tb=short 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -70) > gpurun_out/r2_test3.log; tail -50 gpurun_out/r2_test3.log
timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 256 197 12 0 64 577 16 0 > gpurun_out/r2_attn_bench_tc3.txt 2>&1; cat gpurun_out/r2_attn_bench_tc3.txt
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:attention_tc_bwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_bwd50 python tools/attn_bench.py 1024 50 12 0 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:attention_tc_fwd -s 3 -c 1 -o gpurun_out/prof_attn_fwd50 python tools/attn_bench.py 1024 50 12 0 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:attention_tc_fwd -s 3 -c 1 -o gpurun_out/prof_attn_fwd577 python tools/attn_bench.py 64 577 16 0 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_gelugrad2 python tools/one_gemm.py gelugrad > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_mulaux2 python tools/one_gemm.py mulaux > /dev/null 2>&1
ls gpurun_out/*.ncu-rep | wc -l
