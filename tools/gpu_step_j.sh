#!/bin/bash
# 1-GPU check of the 8-bit saved GELU derivative + split-K model: GEMM / model parity tests, ncu of both GELU GEMMs and the
# in_proj weight gradient, short bench with the per-signature table
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -x -q -m gpu --tb=short 2>&1 | grep -v "UserWarning\|Consider using\|^$" | tail -25) > gpurun_out/r2_test7.log; tail -12 gpurun_out/r2_test7.log
for w in gelugrad mulaux wgradqkv qkv; do timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_${w}4 python tools/one_gemm.py $w > /dev/null 2>&1; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench5.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline_gemm_family']['frac'], d['parity']['ok'], d['hbm_peak_allocated_gb'])
        for s in d['roofline_gemm_signatures']: print(s['epilogue'], s['m'], s['n'], s['k'], round(s['share_of_step'],4), round(s['frac'],3))
PY
tail -2 gpurun_out/r2_bench5.err
