#!/bin/bash
# One `ncu --set full` capture per hot kernel (B200_PROFILING.md recipe), bench at local batch 1024, 1 GPU.
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 \
    python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc $?"
}
cap attn_bwd 'attention_bwd_kernel' 30
cap attn_fwd 'attention_fwd_kernel' 30
cap gemm_fc 'gemm_tc_kernelILi256ELi1' 14
cap gemm_dgelu 'gemm_tc_kernelILi256ELi3' 14
cap gemm_qkv 'gemm_tc_kernelILi256ELi0' 40
cap ln_bwd 'layernorm_bwd_kernel' 30
ls -la gpurun_out/*.ncu-rep
