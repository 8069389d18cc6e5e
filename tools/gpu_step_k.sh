#!/bin/bash
# 1-GPU check after the single-tile attention backward barrier fix and the centred d(logit_scale) sum
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_loss.py tests/test_gpu_model.py -x -q -m gpu --tb=short 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -12) > gpurun_out/r2_test8.log; tail -6 gpurun_out/r2_test8.log
timeout 300 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1 2048 77 12 1 > gpurun_out/r2_attn_bench_tc6.txt 2>&1; cat gpurun_out/r2_attn_bench_tc6.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench6.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline_gemm_family']['frac'], d['parity'])
PY
tail -2 gpurun_out/r2_bench6.err
