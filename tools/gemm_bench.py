"""Micro-benchmark of the GEMM family on the train-step shapes (CUDA events, warm cache excluded by cycling
buffers > L2).  Usage: python tools/gemm_bench.py [batch]   (CLIPN_GEMM_PAIR=0 to force the single-CTA kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_clip_b200 import _lib as L  # noqa: E402
from open_clip_b200 import ops  # noqa: E402

BF16 = torch.bfloat16


def bench(name, fn, flops, iters=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:44s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s")
    return ms


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = "cuda"
    print("pair kernel:", os.environ.get("CLIPN_GEMM_PAIR", "1"), "batch", B)
    total = 0.0
    for tower, M, d in (("vision", B * 50, 768), ("text", B * 77, 512)):
        x = torch.randn(M, d, device=dev).to(BF16)
        x4 = torch.randn(M, 4 * d, device=dev).to(BF16)
        x3 = torch.randn(M, 3 * d, device=dev).to(BF16)
        wqkv = torch.randn(3 * d, d, device=dev).to(BF16) * 0.02
        wo = torch.randn(d, d, device=dev).to(BF16) * 0.02
        wfc = torch.randn(4 * d, d, device=dev).to(BF16) * 0.02
        wpr = torch.randn(d, 4 * d, device=dev).to(BF16) * 0.02
        b3, b1, b4 = (torch.zeros(n, device=dev, dtype=BF16) for n in (3 * d, d, 4 * d))
        o3, o1, o4, o4b = torch.empty_like(x3), torch.empty_like(x), torch.empty_like(x4), torch.empty_like(x4)
        g_qkv = torch.zeros(3 * d, d, device=dev)
        g_fc = torch.zeros(4 * d, d, device=dev)
        g_pr = torch.zeros(d, 4 * d, device=dev)
        g_o = torch.zeros(d, d, device=dev)
        f = lambda n, k: 2.0 * M * n * k
        t = 0.0
        t += bench(f"{tower} qkv fwd  [M,{3*d},{d}] STORE+bias", lambda: ops.gemm(x, wqkv, bias=b3, out=o3), f(3 * d, d))
        t += bench(f"{tower} out fwd  [M,{d},{d}] RESID", lambda: ops.gemm(x, wo, bias=b1, aux=x, epilogue=L.EPI_BIAS_RESID, out=o1), f(d, d))
        t += bench(f"{tower} fc  fwd  [M,{4*d},{d}] GELU", lambda: ops.gemm(x, wfc, bias=b4, epilogue=L.EPI_BIAS_GELU, out=o4, out2=o4b), f(4 * d, d))
        t += bench(f"{tower} proj fwd [M,{d},{4*d}] RESID", lambda: ops.gemm(x4, wpr, bias=b1, aux=x, epilogue=L.EPI_BIAS_RESID, out=o1), f(d, 4 * d))
        t += bench(f"{tower} proj dgrad [M,{4*d},{d}] DGELU", lambda: ops.gemm(x, wpr, b_mn=True, epilogue=L.EPI_DGELU, aux=x4, out=o4, out2=o4b), f(4 * d, d))
        t += bench(f"{tower} fc dgrad [M,{d},{4*d}]", lambda: ops.gemm(x4, wfc, b_mn=True, out=o1), f(d, 4 * d))
        t += bench(f"{tower} out dgrad [M,{d},{d}]", lambda: ops.gemm(x, wo, b_mn=True, out=o1), f(d, d))
        t += bench(f"{tower} qkv dgrad [M,{d},{3*d}]", lambda: ops.gemm(x3, wqkv, b_mn=True, out=o1), f(d, 3 * d))
        for nm, dy, xx, gw in (("proj", x, x4, g_pr), ("fc", x4, x, g_fc), ("out", x, x, g_o), ("qkv", x3, x, g_qkv)):
            s = ops.wgrad_splits(gw.shape[0], gw.shape[1], M)
            t += bench(f"{tower} {nm} wgrad [{gw.shape[0]},{gw.shape[1]},M] splits={s}",
                       lambda: ops.gemm(dy, xx, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw, splits=s),
                       2.0 * M * gw.shape[0] * gw.shape[1])
        print(f"--- {tower}: one block's GEMMs {t:.3f} ms -> x12 = {12 * t:.1f} ms")
        if os.environ.get("CLIPN_BENCH_DIAG"):
            # epilogue / store-path diagnostics on the c_fc shape: one vs two outputs, with and without a mainloop
            xs = x[:, :64].contiguous()
            ws = wfc[:, :64].contiguous()
            bench(f"{tower} DIAG fc shape STORE (1 output)", lambda: ops.gemm(x, wfc, bias=b4, out=o4), f(4 * d, d))
            bench(f"{tower} DIAG fc shape GELU K=64", lambda: ops.gemm(xs, ws, bias=b4, epilogue=L.EPI_BIAS_GELU, out=o4, out2=o4b), f(4 * d, 64))
            bench(f"{tower} DIAG fc shape STORE K=64", lambda: ops.gemm(xs, ws, bias=b4, out=o4), f(4 * d, 64))
            bench(f"{tower} DIAG fc shape DGELU K=64", lambda: ops.gemm(xs, wpr[:64].contiguous(), b_mn=True, epilogue=L.EPI_DGELU, aux=x4, out=o4, out2=o4b), f(4 * d, 64))
            bench(f"{tower} DIAG torch copy 2x[M,4d] (write-heavy HBM reference)", lambda: (o4.copy_(x4), o4b.copy_(x4)), 1.0)
        total += 12 * t
        del x, x4, x3, o3, o1, o4, o4b
    print(f"=== all tower GEMMs per step: {total:.1f} ms")


if __name__ == "__main__":
    main()
