"""Peer-read bandwidth over NVLink as THIS repo uses it (run under torchrun, >= 2 ranks):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_bench.py
Times, per rank, for the ViT-B-32 feature exchange (B=4096, E=512: 4 MB per block, 2 blocks per rank):
  (a) torch copy_ from the peer-mapped symmetric buffer (cudaMemcpy D2D peer, copy engines)
  (b) libclipn's gather kernel (coalesced 16-byte loads from all SMs)
  (c) NCCL all_gather_into_tensor of the same bytes (the reference's gather_features)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_clip_b200 import comm, ops  # noqa: E402


def timeit(fn, sync, iters=10):
    for _ in range(3):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    B, E = 4096, 512
    g = comm.FeatureGather(B, E, dev)
    assert g.mode == "peer", g.why_nccl
    img = torch.randn(B, E, device=dev).to(torch.bfloat16)
    txt = torch.randn(B, E, device=dev).to(torch.bfloat16)
    _, _, img_ptrs, txt_ptrs = g.publish(img, txt)

    def sync():
        torch.cuda.synchronize()
        dist.barrier()

    sync()
    remote_mb = (world - 1) * 2 * B * E * 2 / 1e6
    total_mb = world * 2 * B * E * 2 / 1e6
    peers = [g.hdl.get_buffer(r, (2, 2, B, E), torch.bfloat16) for r in range(world)]

    def ce_copy():
        for r in range(world):
            g.all_img[r * B:(r + 1) * B].copy_(peers[r][0, 0])
            g.all_txt[r * B:(r + 1) * B].copy_(peers[r][0, 1])
    us = timeit(ce_copy, sync)
    print(f"[rank {rank}] torch peer copy_ x{2 * world}: {us:8.1f} us  {remote_mb / us * 1e3:7.1f} GB/s remote ({total_mb:.1f} MB total)")
    us = timeit(lambda: ops.peer_gather(txt_ptrs, img_ptrs, B, E, g.all_txt, g.all_img), sync)
    ok = bool(torch.equal(g.all_img[rank * B:(rank + 1) * B], img)) and bool(torch.equal(g.all_txt[rank * B:(rank + 1) * B], txt))
    print(f"[rank {rank}] clipn_peer_gather:        {us:8.1f} us  {remote_mb / us * 1e3:7.1f} GB/s remote  local block ok={ok}")
    us = timeit(lambda: (dist.all_gather_into_tensor(g.all_img, img), dist.all_gather_into_tensor(g.all_txt, txt)), sync)
    print(f"[rank {rank}] nccl all_gather x2:       {us:8.1f} us  {remote_mb / us * 1e3:7.1f} GB/s remote")
    scale = torch.tensor([14.28], device=dev)
    us = timeit(lambda: ops.clip_fwd_fused(img, txt, txt_ptrs, img_ptrs, rank, scale, g.all_txt, g.all_img), sync)
    print(f"[rank {rank}] clip_fwd_fused (gather + GEMM + combine): {us:8.1f} us")
    one_i, one_t = [g.all_img.data_ptr()], [g.all_txt.data_ptr()]
    big_i, big_t = g.all_img, g.all_txt
    us = timeit(lambda: (ops.clip_lse_fwd(img, big_t, scale, rank * B), ops.clip_lse_fwd(txt, big_i, scale, rank * B)), sync)
    print(f"[rank {rank}] generic lse fwd x2 on the local gathered operands: {us:8.1f} us")
    # ---- the same call after ~60 ms without any NVLink traffic (as in a training step: the towers run in between) ----
    # GPU kept busy by local GEMMs so the host runs ahead; stage times from the library's own events.
    a = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    w = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    tiny_t, tiny_i = torch.empty(world * 8, E, dtype=torch.bfloat16, device=dev), torch.empty(world * 8, E, dtype=torch.bfloat16, device=dev)
    for warm in (False, True):
        tot = []
        ops.stage_timing(True)
        for k in range(6):
            for _ in range(80):
                torch.mm(a, w)
            g.publish(img, txt)
            if warm:   # a 8-row peer read ahead of the real one: does waking the links earlier help?
                ops.peer_gather(txt_ptrs, img_ptrs, 8, E, tiny_t, tiny_i)
                torch.mm(a, w)
            _, _, ip, tp = g.publish(img, txt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.clip_fwd_fused(img, txt, tp, ip, rank, scale, g.all_txt, g.all_img)
            e1.record()
            torch.cuda.synchronize()
            tot.append(e0.elapsed_time(e1) * 1e3)
        n, tg, tm, tc = ops.stage_times()
        ops.stage_timing(False)
        print(f"[rank {rank}] fused forward after 60 ms of link idle (warm-up read {warm}): total {sorted(tot)[len(tot) // 2]:.1f} us "
              f"(median of {len(tot)}); stages over {n} calls: gather {tg * 1e3:.1f} GEMM {tm * 1e3:.1f} combine {tc * 1e3:.1f} us")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
