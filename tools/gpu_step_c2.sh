#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_optim.py "tests/test_gpu_model.py::test_vitl14_336_full_depth_vs_reference_fixture" tests/test_gpu_model.py::test_accum_freq_feature_cache_algorithm_on_native_objects tests/test_gpu_elementwise.py -q -m gpu -s 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -60) > gpurun_out/r2_attn_test2.log; tail -45 gpurun_out/r2_attn_test2.log
timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 256 197 12 0 64 577 16 0 > gpurun_out/r2_attn_bench_tc2.txt 2>&1; cat gpurun_out/r2_attn_bench_tc2.txt
