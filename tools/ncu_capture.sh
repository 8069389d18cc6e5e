#!/bin/bash
# One `ncu --set full` capture per hot kernel (B200_PROFILING.md recipe), bench at local batch 1024, 1 GPU.
# usage: tools/ncu_capture.sh name:regex:skip ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read name regex skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -f -o gpurun_out/prof_$name \
    python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc $? $(ls -la gpurun_out/prof_$name.ncu-rep 2>/dev/null | wc -l)"
done
