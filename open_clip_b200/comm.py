"""Peer-memory feature exchange over NVLink 5 / NVSwitch.

Replaces the two NCCL all-gathers of `gather_features` (reference loss.py:29-54): every rank writes its
[B,E] bf16 image/text features into a symmetric (peer-mapped) buffer once; the logits GEMM then reads every
peer's buffer directly through per-rank TMA tensor maps (libclipn `feats_cols`), so the gather is fused
into the GEMM's operand loads.  The backward's [N,E] gradient reduce-scatter is eliminated by exchanging
only the two row-LSE vectors (2*N fp32), see loss.py in this package.

torch.distributed supplies the bootstrap (rendezvous + symmetric-memory handles), as in the reference
(open_clip_train/distributed.py:102-166); no NCCL collective touches the feature matrices.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class PeerFeatureExchange:
    """Symmetric [2, B, E] bf16 buffer (slot 0 image, slot 1 text) mapped into every rank."""

    def __init__(self, batch: int, embed: int, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.batch, self.embed = batch, embed
        self.buf = symm_mem.empty((2, batch, embed), dtype=torch.bfloat16, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        slot = batch * embed * 2
        self.img_ptrs: List[int] = [int(p) for p in self.hdl.buffer_ptrs]
        self.txt_ptrs: List[int] = [int(p) + slot for p in self.hdl.buffer_ptrs]

    def publish(self, image_features: torch.Tensor, text_features: torch.Tensor):
        """barrier (peers finished reading last step's data) -> write -> barrier (data visible to peers)."""
        self.hdl.barrier(channel=0)
        self.buf[0].copy_(image_features)
        self.buf[1].copy_(text_features)
        self.hdl.barrier(channel=1)

    def local_image(self) -> torch.Tensor:
        return self.buf[0]

    def local_text(self) -> torch.Tensor:
        return self.buf[1]


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
