"""Time the attention kernels alone: python tools/attn_bench.py B L H [causal] [B L H causal ...]  (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_clip_b200 import ops  # noqa: E402


def main():
    args = sys.argv[1:]
    if len(args) > 4:  # several configs in one process: groups of four "B L H causal"
        for i in range(0, len(args), 4):
            run(int(args[i]), int(args[i + 1]), int(args[i + 2]), args[i + 3] == "1")
        return
    run(int(args[0]), int(args[1]), int(args[2]), len(args) > 3 and args[3] == "1")


def run(B, Lq, H, causal):
    d = H * 64
    qkv = (torch.randn(B * Lq, 3 * d, device="cuda") * 0.5).to(torch.bfloat16)
    do = (torch.randn(B * Lq, d, device="cuda") * 0.1).to(torch.bfloat16)
    dbias = torch.zeros(3 * d, device="cuda")
    o, lse = ops.attention_fwd(qkv, B, Lq, H, causal)
    dq = torch.empty_like(qkv)

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    tf = timed(lambda: ops.attention_fwd(qkv, B, Lq, H, causal, out=o))
    tb = timed(lambda: ops.attention_bwd(qkv, o, do, lse, B, Lq, H, causal, out=dq, dbias=dbias))
    flops_f = 4.0 * B * H * Lq * Lq * 64 * (0.5 if causal else 1.0)
    bytes_f = B * Lq * d * 2 * 4
    bytes_b = B * Lq * d * 2 * 8
    print(f"attention B={B} L={Lq} H={H} causal={causal}: fwd {tf*1e3:.0f} us ({flops_f/tf/1e9:.1f} TF/s, "
          f"{bytes_f/tf/1e6:.0f} GB/s algorithmic)  bwd {tb*1e3:.0f} us ({2.5*flops_f/tb/1e9:.1f} TF/s, "
          f"{bytes_b/tb/1e6:.0f} GB/s algorithmic)")


if __name__ == "__main__":
    main()
