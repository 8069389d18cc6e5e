#!/bin/bash
# One GPU trip that measures every switch prepared at the end of round 1 (each line is one python process; a fresh
# box pays ~25 s of `import torch` per process, so configurations are batched per process where the tool allows it).
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/experiments_r02.sh'
mkdir -p gpurun_out
export PYTHONPATH=.
{
  echo "== HBM: write-only vs copy vs read-only (is 3.2 TB/s on the GEMM store path a DRAM limit?)"
  timeout 120 python tools/hbm_write_bench.py 630
  echo "== attention backward with K/V prefetch: parity first"
  CLIPN_ATTN_BWD_PREFETCH=1 timeout 200 python -m pytest tests/test_gpu_attention.py -x -q -m gpu 2>&1 | tail -3
  echo "== attention timing: default, then prefetch"
  timeout 120 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1
  CLIPN_ATTN_BWD_PREFETCH=1 timeout 120 python tools/attn_bench.py 1024 50 12 0 1024 77 8 1
  echo "== towers on two streams: parity, then the step time against the default"
  CLIPN_TOWER_STREAMS=1 timeout 200 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiny or train_steps" 2>&1 | tail -3
  timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-220
  CLIPN_TOWER_STREAMS=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-220
  echo "== GEMM shapes: default tile, then 128-wide tile"
  timeout 200 python tools/gemm_bench.py 1024
  CLIPN_GEMM_TILE_N=128 timeout 200 python tools/gemm_bench.py 1024
} > gpurun_out/experiments_r02.log 2>&1
tail -80 gpurun_out/experiments_r02.log
