#!/bin/bash
# Run the GPU parity suite file-by-file (separate processes: a device trap in one file cannot poison the rest).
# Usage (on the GPU box, via gpurun): bash tools/gpu_round.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
FILES=${@:-"tests/test_gpu_elementwise.py tests/test_gpu_attention.py tests/test_gpu_gemm.py tests/test_gpu_loss.py tests/test_gpu_model.py"}
for f in $FILES; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 240 -x --maxfail=8 > gpurun_out/$name.log 2>&1
  echo "=== $f exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|error" gpurun_out/$name.log | tail -2 | tee -a gpurun_out/summary.txt
done
for f in $FILES; do
  name=$(basename $f .py)
  echo "----- $name (failures)"; grep -E "^(FAILED|ERROR)|Error|assert|clipn:" gpurun_out/$name.log | head -30
done
