"""Time LayerNorm fwd/bwd alone: python tools/ln_bench.py ROWS D   (CUDA events; tensors far larger than L2)."""
import sys

import torch

from open_clip_b200 import ops


def main():
    rows, d = int(sys.argv[1]), int(sys.argv[2])
    x = torch.randn(rows, d, device="cuda").to(torch.bfloat16)
    dy = torch.randn(rows, d, device="cuda").to(torch.bfloat16)
    res = torch.randn(rows, d, device="cuda").to(torch.bfloat16)
    w = torch.ones(d, device="cuda")
    b = torch.zeros(d, device="cuda")
    gw, gb = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    _, mean, rstd = ops.layernorm_fwd(x, w, b, out=y, save_stats=True)

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return a.elapsed_time(e) / n

    tf = timed(lambda: ops.layernorm_fwd(x, w, b, out=y, save_stats=True))
    tb = timed(lambda: ops.layernorm_bwd(dy, x, mean, rstd, w, gw, gb, resid=res, out=dx))
    bf, bb = rows * d * 2 * 2, rows * d * 2 * 4
    print(f"layernorm rows={rows} d={d}: fwd {tf*1e3:.1f} us ({bf/tf/1e6:.0f} GB/s)  bwd+resid {tb*1e3:.1f} us ({bb/tb/1e6:.0f} GB/s)")


if __name__ == "__main__":
    main()
