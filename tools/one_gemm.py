"""Launch one GEMM signature a few times (for ncu captures): python tools/one_gemm.py {proj|dgelu|gelu}"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from open_clip_b200 import ops, _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "proj"
M, d = 51200, 768
bf = torch.bfloat16
x = torch.randn(M, d, device="cuda").to(bf)
x4 = torch.randn(M, 4 * d, device="cuda").to(bf)
o4, o4b, o1 = torch.empty_like(x4), torch.empty_like(x4), torch.empty_like(x)
wfc = (torch.randn(4 * d, d, device="cuda") * 0.02).to(bf)
wpr = (torch.randn(d, 4 * d, device="cuda") * 0.02).to(bf)
b4 = torch.zeros(4 * d, device="cuda", dtype=bf)
cs = torch.zeros(4 * d, device="cuda")
for _ in range(3):
    if which == "proj":
        ops.gemm(x4, wpr, out=o1)
    elif which == "dgelu":
        ops.gemm(x, wpr, b_mn=True, epilogue=L.EPI_DGELU, aux=x4, out=o4, col_sum=cs)
    elif which == "gelu":
        ops.gemm(x, wfc, bias=b4, epilogue=L.EPI_BIAS_GELU, out=o4, out2=o4b)
    elif which == "wgrad":     # split-K weight gradient (both operands MN-major, TMA reduce-add)
        gw = torch.zeros(4 * d, d, device="cuda")
        ops.gemm(x4, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw, splits=ops.wgrad_splits(4 * d, d, M))
    elif which == "wgradqkv":  # in_proj weight gradient at the benchmarked size: [2304, 768] = dqkv^T x, K = 204800
        Mq = 204800
        if _ == 0:
            xq = torch.randn(Mq, d, device="cuda").to(bf)
            dq = torch.randn(Mq, 3 * d, device="cuda").to(bf)
        gw = torch.zeros(3 * d, d, device="cuda")
        ops.gemm(dq, xq, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw, splits=ops.wgrad_splits(3 * d, d, Mq))
    elif which == "qkv":       # in_proj forward (plain store epilogue), K = 768
        w3 = (torch.randn(3 * d, d, device="cuda") * 0.02).to(bf)
        o3 = torch.empty(M, 3 * d, device="cuda", dtype=bf)
        ops.gemm(x, w3, out=o3)
    elif which == "resid":     # out_proj / c_proj forward with the residual add
        b1 = torch.zeros(d, device="cuda", dtype=bf)
        ops.gemm(x4, wpr, bias=b1, aux=x, epilogue=L.EPI_BIAS_RESID, out=o1)
    elif which == "gelugrad":  # forward c_fc GEMM of the default (fast) activation mode: writes gelu'(h) and gelu(h)
        ops.gemm(x, wfc, bias=b4, epilogue=L.EPI_BIAS_GELU_GRAD, out=o4, out2=o4b)
    elif which == "mulaux":    # backward c_proj dgrad x saved gelu'(h) + fused c_fc bias gradient
        ops.gemm(x, wpr, b_mn=True, epilogue=L.EPI_MUL_AUX, aux=x4, out=o4, col_sum=cs)
torch.cuda.synchronize()
