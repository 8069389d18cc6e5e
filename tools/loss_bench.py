"""CUDA-event timings of the contrastive-loss kernels on ONE GPU (python tools/loss_bench.py [B] [E]).

World sizes > 1 are emulated by listing W local [B,E] buffers as the per-rank column pointers: same kernel, same
tile walk, same N = W*B, only the bytes come from local HBM instead of NVLink — this isolates the kernel's tensor /
epilogue efficiency from the fabric.  Every timed result is checked against fp32 torch first.
"""
import sys
import os

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_clip_b200 import ops  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
PEAK = 1461.6  # TFLOP/s sustained (MEASURED_PEAKS.json)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()  # evict L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # median, us


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    torch.manual_seed(0)
    scale = torch.tensor([14.2857], device="cuda")
    for W in (1, 2, 8):
        N = W * B
        imgs = [F.normalize(torch.randn(B, E, device="cuda"), dim=-1).to(BF16) for _ in range(W)]
        txts = [F.normalize(torch.randn(B, E, device="cuda"), dim=-1).to(BF16) for _ in range(W)]
        g_img = torch.zeros(N, E, dtype=BF16, device="cuda") if W > 1 else None
        g_txt = torch.zeros(N, E, dtype=BF16, device="cuda") if W > 1 else None
        rank = W // 2
        ip, tp = [t.data_ptr() for t in imgs], [t.data_ptr() for t in txts]
        lse, loss = ops.clip_fwd_fused(imgs[rank], txts[rank], tp, ip, rank, scale, g_txt, g_img)
        torch.cuda.synchronize()
        all_i, all_t = torch.cat(imgs), torch.cat(txts)
        s1 = scale * imgs[rank].float() @ all_t.float().T
        s2 = scale * txts[rank].float() @ all_i.float().T
        ref = torch.stack([torch.logsumexp(s1, 1), torch.logsumexp(s2, 1)])
        idx = torch.arange(B, device="cuda")
        off = rank * B if W > 1 else 0
        ref_loss = ((ref[0] - s1[idx, off + idx]).mean() + (ref[1] - s2[idx, off + idx]).mean()) / 2
        err = float((lse - ref).abs().max())
        gerr = 0.0
        if W > 1:
            gerr = float((g_txt.float() - all_t.float()).abs().max() + (g_img.float() - all_i.float()).abs().max())
        print(f"W={W} N={N}: fused fwd max|lse err| {err:.2e}  loss {float(loss):.5f} vs {float(ref_loss):.5f}  "
              f"gather err {gerr:.1e}")
        assert err < 2e-3 and abs(float(loss) - float(ref_loss)) < 2e-3 and gerr == 0.0
        del s1, s2
        fl = 2 * 2.0 * B * N * E
        us = timeit(lambda: ops.clip_fwd_fused(imgs[rank], txts[rank], tp, ip, rank, scale, g_txt, g_img))
        print(f"  fused fwd (2 dirs + combine)      {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  {fl / us / 1e6 / PEAK:5.1%} of sustained peak")
        ai, at = (all_i, all_t) if W > 1 else (imgs[0], txts[0])
        us = timeit(lambda: (ops.clip_lse_fwd(imgs[rank], at, scale, off), ops.clip_lse_fwd(txts[rank], ai, scale, off)))
        print(f"  generic lse fwd x2 (local cols)   {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  {fl / us / 1e6 / PEAK:5.1%}")
        col = torch.cat([lse[1]] * W) if W > 1 else lse[1]
        acc = torch.zeros(4, device="cuda")
        f1 = 2.0 * B * N * E
        dl = ops.clip_dlogits(imgs[rank], at, scale, off, lse[0], col, 1.0, 1.0 / (2 * B), acc[0:2])
        us = timeit(lambda: ops.clip_dlogits(imgs[rank], at, scale, off, lse[0], col, 1.0, 1.0 / (2 * B), acc[0:2]))
        print(f"  dlogits (1 dir)                   {us:8.1f} us  {f1 / us / 1e6:7.1f} TF/s  {f1 / us / 1e6 / PEAK:5.1%}")
        d = ops.clip_dfeat(dl, at, scale)
        refd = scale * dl.float() @ at.float()  # dl: centred softmax parts (mean + one-hot are the caller's fp32 init)
        rel = float((d - refd).norm() / refd.norm())
        assert rel < 2e-3, rel
        us = timeit(lambda: ops.clip_dfeat(dl, at, scale))
        print(f"  dfeat split-K (1 dir, splits {ops.wgrad_splits(B, E, N)})     {us:8.1f} us  {f1 / us / 1e6:7.1f} TF/s  {f1 / us / 1e6 / PEAK:5.1%}  rel {rel:.1e}")
        del dl, d, refd


if __name__ == "__main__":
    main()
