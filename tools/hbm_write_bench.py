"""Is the 3.2 TB/s seen on the GEMM store path an HBM write ceiling or a store-path limit?
Times write-only (fill), read-only (sum) and copy streams over buffers far larger than the 126 MB L2 with CUDA events.
    python tools/hbm_write_bench.py [MB]"""
import sys

import torch


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
    return a.elapsed_time(b) / n * 1e-3


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 630
    n = mb * 1000 * 1000 // 2
    x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    x.normal_()
    t_fill = timed(lambda: y.fill_(1.0))
    t_zero = timed(lambda: y.zero_())
    t_copy = timed(lambda: y.copy_(x))
    t_read = timed(lambda: x.view(torch.int16).max())
    gb = n * 2 / 1e9
    print(f"buffer {gb:.2f} GB: fill {gb / t_fill:.0f} GB/s write-only, zero_ {gb / t_zero:.0f} GB/s write-only, "
          f"copy {2 * gb / t_copy:.0f} GB/s total ({gb / t_copy:.0f} written), max-reduce {gb / t_read:.0f} GB/s read-only")


if __name__ == "__main__":
    main()
