#!/bin/bash
# 1-GPU check of the 16-warp GELU epilogues: GEMM parity tests, micro-benchmark, ncu of the two signatures, short bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_loss.py -x -q -m gpu --tb=short 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -12) > gpurun_out/r2_test6.log; tail -5 gpurun_out/r2_test6.log
timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/r2_gemmbench_w16.txt 2>&1; grep -i "fc  fwd\|gelu\|mul\|one block\|all tower" gpurun_out/r2_gemmbench_w16.txt | head -20
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_gelugrad3 python tools/one_gemm.py gelugrad > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_mulaux3 python tools/one_gemm.py mulaux > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench4.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_gemm_family']['frac'], d['roofline_logits_gemm'], d['parity']['ok'])
        for s in d['roofline_gemm_signatures']: print(s['epilogue'], s['m'], s['n'], s['k'], round(s['share_of_step'],4), round(s['frac'],3))
PY
tail -2 gpurun_out/r2_bench4.err
