"""GPU parity, world_size 2 (needs 2 GPUs; skipped otherwise): the peer-memory fused ClipLoss / SigLipLoss vs the
oracle's process-group-free restatement of the reference's multi-rank semantics (pinned to a real gloo run of the
reference by oracle/gen_golden.py and tests/test_oracle.py) for all four (local_loss, gather_with_grad) modes."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

WORLD, B, E = 2, 256, 64


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from open_clip_b200.loss import NativeClipLoss, NativeSigLipLoss
    from oracle import clip_oracle as O

    g = torch.Generator().manual_seed(11)
    img = [F.normalize(torch.randn(B, E, generator=g), dim=-1).to(torch.bfloat16) for _ in range(world)]
    txt = [F.normalize(torch.randn(B, E, generator=g), dim=-1).to(torch.bfloat16) for _ in range(world)]
    scale = torch.tensor(14.2857)

    def rel(a, b):
        return float((a.float().cpu() - b.float()).norm() / (b.float().norm() + 1e-12))

    for local_loss in (True, False):
        for gwg in (True, False):
            gi = img[rank].cuda().clone().requires_grad_(True)
            gt = txt[rank].cuda().clone().requires_grad_(True)
            gs = scale.cuda().clone().requires_grad_(True)
            loss = NativeClipLoss(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=world)(gi, gt, gs)
            loss.backward()
            torch.cuda.synchronize()
            ri = [t.float().clone().requires_grad_(True) for t in img]
            rt = [t.float().clone().requires_grad_(True) for t in txt]
            rs = scale.clone().requires_grad_(True)
            ref = O.clip_loss_ranks(ri, rt, rs, local_loss, gwg)
            # feature grads: what autograd + the (reduce-scatter / splice) gather gives rank r = sum over ranks' losses
            sum(ref).backward()
            tag = (local_loss, gwg, rank)
            assert abs(float(loss) - float(ref[rank])) < 3e-2, (tag, float(loss), float(ref[rank]))
            assert rel(gi.grad, ri[rank].grad) < 1.5e-2, (tag, "d_img", rel(gi.grad, ri[rank].grad))
            assert rel(gt.grad, rt[rank].grad) < 1.5e-2, (tag, "d_txt", rel(gt.grad, rt[rank].grad))
            # logit_scale: each rank differentiates ITS loss only
            rs2 = scale.clone().requires_grad_(True)
            O.clip_loss_ranks([t.float() for t in img], [t.float() for t in txt], rs2, local_loss, gwg)[rank].backward()
            assert abs(float(gs.grad) - float(rs2.grad)) < 2e-2 * abs(float(rs2.grad)) + 1e-5, (tag, "d_scale")

    # SigLIP (all dist_impl's are the same sum; ours reads peer blocks in place)
    bias = torch.tensor(-10.0)
    gi = img[rank].cuda().clone().requires_grad_(True)
    gt = txt[rank].cuda().clone().requires_grad_(True)
    gs, gb = torch.tensor(10.0).cuda().requires_grad_(True), bias.cuda().clone().requires_grad_(True)
    loss = NativeSigLipLoss(rank=rank, world_size=world)(gi, gt, gs, gb)
    loss.backward()
    torch.cuda.synchronize()
    ri = [t.float().clone().requires_grad_(True) for t in img]
    rt = [t.float().clone().requires_grad_(True) for t in txt]
    ref = O.siglip_loss_ranks(ri, rt, torch.tensor(10.0), bias)
    sum(ref).backward()
    assert abs(float(loss) - float(ref[rank])) < 2e-2 * abs(float(ref[rank])) + 1e-3
    assert rel(gi.grad, ri[rank].grad) < 2e-2 and rel(gt.grad, rt[rank].grad) < 2e-2
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_losses_match_reference_semantics():
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(WORLD, 29741), nprocs=WORLD, join=True)
