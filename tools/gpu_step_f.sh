#!/bin/bash
# BASELINE configs 4 and 5 on ONE GPU (smoke + timing before the 8-GPU run)
mkdir -p gpurun_out
timeout 900 python bench.py --model ViT-L-14-336 --batch 2048 --grad-checkpointing --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c4_1gpu.json 2> gpurun_out/r2_c4_1gpu.err; tail -c 1500 gpurun_out/r2_c4_1gpu.json; tail -3 gpurun_out/r2_c4_1gpu.err
timeout 600 python bench.py --model ViT-B-16 --siglip --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c5_1gpu.json 2> gpurun_out/r2_c5_1gpu.err; tail -c 1500 gpurun_out/r2_c5_1gpu.json; tail -3 gpurun_out/r2_c5_1gpu.err
