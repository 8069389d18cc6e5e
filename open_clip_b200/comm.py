"""Peer-memory feature exchange over NVLink 5 / NVSwitch.

Replaces the two NCCL all-gathers of `gather_features` (reference loss.py:29-54): every rank writes its
[B,E] bf16 image/text features into a symmetric (peer-mapped) buffer once; the fused logits kernel then reads
every peer's buffer directly (libclipn `txt_cols` / `img_cols`: coalesced P2P loads by all SMs, every byte crossing
NVLink once at the fabric rate; CLIPN_PEER_DIRECT=1 instead streams TMA tiles from the peers inside the GEMM, which is
latency-bound per SM — see csrc/loss.cu) and leaves a local gathered copy behind for the backward.  The backward reads only that local copy; the [N,E] gradient reduce-scatter of
`_all_gather_with_grad` (loss.py:23-26) is eliminated by exchanging only the two row-LSE vectors (2*N fp32) for
ClipLoss and nothing at all for SigLipLoss (see loss.py in this package).

Envelope of the peer path: one NVLink domain with `torch.distributed._symmetric_memory`, world <= 8, embed dim a
multiple of 64 and <= 1024.  Anything else — more ranks, several nodes — takes the NCCL fallback: `all_gather_into_tensor` into the same local gathered
buffers, followed by the generic single-map kernels.  Same values, same gradient conventions.

torch.distributed supplies the bootstrap (rendezvous + symmetric-memory handles), as in the reference
(open_clip_train/distributed.py:102-166).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops

BF16 = torch.bfloat16


class FeatureGather:
    """Per-(batch, embed) exchange state of one loss module.

    peer mode : symmetric [2 slots, 2 (image, text), B, E] bf16 buffer mapped into every rank.  Steps alternate
                between the two slots, so ONE barrier per step is enough: the barrier of step t (data of step t
                visible everywhere) also proves that every rank finished reading step t-1's slot — which is the
                slot step t+1 overwrites.
    nccl mode : no symmetric memory; `gather_nccl` all-gathers into the local buffers.
    Both modes own `all_img` / `all_txt`: local bf16 [W*B, E] copies of every rank's features for the backward.
    """

    def __init__(self, batch: int, embed: int, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.batch, self.embed = batch, embed
        self.all_img = torch.empty((self.world * batch, embed), dtype=BF16, device=device)
        self.all_txt = torch.empty((self.world * batch, embed), dtype=BF16, device=device)
        self.step = 0
        self.mode = "nccl"
        self.why_nccl = ""
        if os.environ.get("CLIPN_FORCE_NCCL_GATHER") == "1":
            self.why_nccl = "CLIPN_FORCE_NCCL_GATHER=1"
        elif ops.peer_gemm_tile_n(1, self.world * batch, embed) == 0:
            self.why_nccl = (f"shape outside the fused kernel's envelope (world {self.world}, batch {batch}, "
                             f"embed {embed})")
        else:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.buf = symm_mem.empty((2, 2, batch, embed), dtype=BF16, device=device)
                self.hdl = symm_mem.rendezvous(self.buf, self.group)
                ptrs = [int(p) for p in self.hdl.buffer_ptrs]
                slot, half = 2 * batch * embed * 2, batch * embed * 2
                self._img_ptrs = [[p + s * slot for p in ptrs] for s in range(2)]
                self._txt_ptrs = [[p + s * slot + half for p in ptrs] for s in range(2)]
                self.mode = "peer"
            except Exception as exc:  # no symmetric memory on this transport (multi-node, no P2P, old torch)
                self.why_nccl = f"symmetric memory unavailable: {type(exc).__name__}: {exc}"

    def publish(self, image_features: torch.Tensor, text_features: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor,
                                                                                           List[int], List[int]]:
        """peer mode: write this step's features into the current slot, barrier; returns (local image view, local text
        view, per-rank image pointers, per-rank text pointers)."""
        s = self.step & 1
        self.step += 1
        self.buf[s, 0].copy_(image_features)
        self.buf[s, 1].copy_(text_features)
        self.hdl.barrier(channel=0)
        return self.buf[s, 0], self.buf[s, 1], self._img_ptrs[s], self._txt_ptrs[s]

    def gather_nccl(self, image_features: torch.Tensor, text_features: torch.Tensor):
        """Fallback: the reference's two all-gathers (loss.py:43-46), into the local gathered buffers."""
        dist.all_gather_into_tensor(self.all_img, image_features.contiguous(), group=self.group)
        dist.all_gather_into_tensor(self.all_txt, text_features.contiguous(), group=self.group)


def all_gather_vectors(vecs: torch.Tensor, group=None) -> torch.Tensor:
    """[k, B] per rank -> [k, W*B] (rank-major along the last dim). 2*N fp32 per step: the only exchange the
    backward needs (replaces the reduce-scatter of [N,E] feature grads, loss.py:23-26). Works on gloo (CPU tests)
    and nccl."""
    world = dist.get_world_size(group)
    k, b = vecs.shape
    out = torch.empty((world * k, b), dtype=vecs.dtype, device=vecs.device)
    dist.all_gather_into_tensor(out, vecs.contiguous(), group=group)
    return out.view(world, k, b).permute(1, 0, 2).reshape(k, world * b).contiguous()


def clip_grad_convention(local_loss: bool, gather_with_grad: bool, batch: int, world: int):
    """(gscale, col_w, global_value) for ClipLoss's four (local_loss, gather_with_grad) modes — the
    gradient-scale conventions of SURVEY §8e, pinned by tests/golden/loss_w*.pt.

    gscale : scale of d(loss)/d(logits) terms in the [B x N] tile kernels
    col_w  : weight of the column-softmax term (other ranks' rows that see my features as columns)
    global_value : loss value is the mean over ranks (global loss) instead of the local one
    """
    n = batch * world
    if world == 1:
        return 1.0 / (2 * batch), 1.0, False
    if local_loss:
        return 1.0 / (2 * batch), (1.0 if gather_with_grad else 0.0), False
    # global [N x N] loss on every rank
    if gather_with_grad:
        return 1.0 / (2 * batch), 1.0, True   # W identical losses back-propagate through the gather: W/(2N)
    return 1.0 / (2 * n), 1.0, True
