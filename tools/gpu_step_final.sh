#!/bin/bash
# final sanity of the tree as committed: smoke() and the default bench line (value + e2e + cpu_baseline + parity)
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench_final.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d['e2e'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_gemm_family']['frac'], d['gpu_launches'], d['clocks'], d['parity']['ok'])
PY
tail -2 gpurun_out/r2_bench_final.err
