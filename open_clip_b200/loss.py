"""NativeClipLoss / NativeSigLipLoss — drop-ins for the reference ClipLoss (loss.py:57-141) and SigLipLoss
(loss.py:314-489): same constructor kwargs, same forward signature, same returned value / dict, same gradient
conventions for every (local_loss, gather_with_grad) combination — computed by libclipn's fused kernels.

Forward: the [B x N] logits are never materialised; a tcgen05 GEMM with an online log-sum-exp epilogue produces
the row LSE and the label logit for both directions, reading every rank's features straight from peer memory
(comm.PeerFeatureExchange) instead of all-gathering them.
Backward: d(logits) tiles are recomputed from the LSE vectors (the only cross-rank exchange is 2*N fp32), written
once in bf16 and contracted with the (peer-resident) features by two more tcgen05 GEMMs.  This reproduces the
reduce-scatter semantics of `gather_with_grad` exactly (each rank's feature gradient = sum over ALL ranks'
losses) without moving any [N,E] gradient across NVLink.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import comm, ops
from ._lib import ClipnError

BF16, F32 = torch.bfloat16, torch.float32


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "NativeClipLoss", image_features, text_features, logit_scale, logit_bias):
        # loss.py:100-116 adds logit_bias to every logit; softmax cross-entropy is invariant to a constant added to a
        # whole row, so the value is unchanged and d loss / d logit_bias = sum(softmax - onehot) = 0.
        ctx.bias_like = logit_bias.detach() if logit_bias is not None else None
        if not image_features.is_cuda:
            raise ClipnError("NativeClipLoss runs on CUDA tensors only; there is no CPU fallback")
        B, E = image_features.shape
        W, rank = module.world_size, module.rank
        img = image_features.detach().to(BF16).contiguous()
        txt = text_features.detach().to(BF16).contiguous()
        scale = logit_scale.detach().to(F32).reshape(1).contiguous()  # stays on the device: no host sync
        if W > 1:
            ex = module._exchange(B, E, img.device)
            ex.publish(img, txt)
            img_ptrs, txt_ptrs = ex.img_ptrs, ex.txt_ptrs
            img, txt = ex.local_image(), ex.local_text()
        else:
            img_ptrs, txt_ptrs = [img.data_ptr()], [txt.data_ptr()]
        off = rank * B if W > 1 else 0
        lse_i, pos_i = ops.clip_lse_fwd(img, txt_ptrs, scale, off)  # logits_per_image rows (loss.py:103)
        lse_t, pos_t = ops.clip_lse_fwd(txt, img_ptrs, scale, off)  # logits_per_text rows  (loss.py:104)
        local = ((lse_i - pos_i).mean() + (lse_t - pos_t).mean()) * 0.5
        gscale, col_w, global_value = comm.clip_grad_convention(module.local_loss, module.gather_with_grad, B, W)
        loss = local
        if global_value:
            loss = local.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            loss = loss / W
        ctx.module, ctx.meta = module, (B, E, W, off, scale, gscale, col_w, global_value)
        ctx.saved = (img, txt, img_ptrs, txt_ptrs, lse_i, lse_t)
        ctx.in_dtypes = (image_features.dtype, text_features.dtype, logit_scale.dtype)
        return loss.to(image_features.dtype) if image_features.dtype != F32 else loss

    @staticmethod
    def backward(ctx, dloss):
        B, E, W, off, scale, gscale, col_w, global_value = ctx.meta
        img, txt, img_ptrs, txt_ptrs, lse_i, lse_t = ctx.saved
        if W > 1 and col_w != 0.0:
            both = comm.all_gather_vectors(torch.stack([lse_i, lse_t]))  # [2, N]
            all_lse_i, all_lse_t = both[0], both[1]
        else:
            all_lse_i, all_lse_t = lse_i, lse_t
        acc = torch.zeros(4, dtype=F32, device=img.device)
        # d logits_per_image[B,N] (rows = my images): row softmax uses my image LSE, column term uses every text's LSE
        dl_i = ops.clip_dlogits(img, txt_ptrs, scale, off, lse_i, all_lse_t if col_w else None, col_w, gscale, acc[0:2])
        d_img = ops.clip_dfeat(dl_i, txt_ptrs, E, scale)
        del dl_i
        dl_t = ops.clip_dlogits(txt, img_ptrs, scale, off, lse_t, all_lse_i if col_w else None, col_w, gscale, acc[2:4])
        d_txt = ops.clip_dfeat(dl_t, img_ptrs, E, scale)
        del dl_t
        # d loss / d logit_scale = sum_{dir} sum (P_row - onehot) * <row, col> / (2B)   (own loss only)
        d_scale = (acc[0] + acc[2]) * ((1.0 / (2 * B)) / gscale)
        if global_value:
            d_scale = d_scale.clone()
            dist.all_reduce(d_scale, op=dist.ReduceOp.SUM)
            d_scale = d_scale / W
        g = dloss.to(F32)
        d_img = (d_img * g.to(d_img.dtype)).to(ctx.in_dtypes[0])
        d_txt = (d_txt * g.to(d_txt.dtype)).to(ctx.in_dtypes[1])
        d_bias = torch.zeros_like(ctx.bias_like) if ctx.bias_like is not None else None
        return None, d_img, d_txt, (d_scale * g).to(ctx.in_dtypes[2]), d_bias


class NativeClipLoss(nn.Module):
    """Reference ctor/forward signature: loss.py:59-72,118-141."""

    def __init__(self, local_loss: bool = False, gather_with_grad: bool = False, cache_labels: bool = False,
                 rank: int = 0, world_size: int = 1):
        super().__init__()
        self.local_loss, self.gather_with_grad, self.cache_labels = local_loss, gather_with_grad, cache_labels
        self.rank, self.world_size = rank, world_size
        self._ex: Optional[comm.PeerFeatureExchange] = None

    def _exchange(self, batch: int, embed: int, device) -> comm.PeerFeatureExchange:
        if self._ex is None or self._ex.batch != batch or self._ex.embed != embed:
            self._ex = comm.PeerFeatureExchange(batch, embed, device)
        return self._ex

    def forward(self, image_features, text_features, logit_scale, logit_bias=None, output_dict: bool = False):
        loss = _ClipLossFn.apply(self, image_features, text_features, logit_scale, logit_bias)
        return {"contrastive_loss": loss} if output_dict else loss


class _SigLipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "NativeSigLipLoss", image_features, text_features, logit_scale, logit_bias):
        if not image_features.is_cuda:
            raise ClipnError("NativeSigLipLoss runs on CUDA tensors only; there is no CPU fallback")
        B, E = image_features.shape
        W, rank = module.world_size, module.rank
        img = image_features.detach().to(BF16).contiguous()
        txt = text_features.detach().to(BF16).contiguous()
        scale = logit_scale.detach().to(F32).reshape(1).contiguous()
        bias = logit_bias.detach().to(F32).reshape(1).contiguous()
        need_grad = any(t.requires_grad for t in (image_features, text_features, logit_scale, logit_bias))
        loss_acc = torch.zeros(1, dtype=F32, device=img.device)
        sacc = torch.zeros(2, dtype=F32, device=img.device)
        gscale = 1.0 / B  # loss.py:366 `.sum() / image_features.shape[0]`
        d_img = torch.zeros((B, E), dtype=F32, device=img.device) if need_grad else None
        d_txt_blocks = []
        if W > 1:
            ex = module._exchange(B, E, img.device)
            ex.publish(img, txt)
            img = ex.local_image()
            peers = [ex.hdl.get_buffer(r, (2, B, E), BF16)[1] for r in range(W)]
        else:
            peers = [txt]
        # every other rank's text block is visited exactly once as a negative-only block (loss.py:410-487: all four
        # dist_impl's are the same sum); NVSwitch peers are uniform, so blocks are read in place instead of ring-passed
        for s in range(W):
            r = (rank + s) % W
            dl = ops.siglip_block(img, peers[r], scale, bias, negative_only=(r != rank), gscale=gscale,
                                  loss_acc=loss_acc, scalar_acc=sacc, want_grad=need_grad)
            if need_grad:
                # d_img += scale * dl @ txt_r ; d_txt_r = scale * dl^T @ img
                ops.gemm(dl, peers[r], b_mn=True, epilogue=ops.L.EPI_ACCUM_F32, out=d_img, alpha_dev=scale)
                d_txt_blocks.append((r, ops.gemm(dl, img, a_mn=True, b_mn=True, epilogue=ops.L.EPI_STORE_F32,
                                                 alpha_dev=scale)))
        ctx.saved = (d_img, d_txt_blocks, sacc)
        ctx.meta = (W, rank, B, E)
        ctx.module = module
        ctx.in_dtypes = tuple(t.dtype for t in (image_features, text_features, logit_scale, logit_bias))
        loss = loss_acc[0]
        return loss.to(image_features.dtype) if image_features.dtype != F32 else loss

    @staticmethod
    def backward(ctx, dloss):
        d_img, d_txt_blocks, sacc = ctx.saved
        W, rank, B, E = ctx.meta
        if W == 1:
            d_txt = d_txt_blocks[0][1]
        else:
            # gradients w.r.t. other ranks' text features flow back to their owners (the reverse exchange of
            # NeighbourExchange.backward, loss.py:287-307): sum over ranks of the [W,B,E] block tensor, keep own slice
            full = torch.zeros((W, B, E), dtype=F32, device=d_img.device)
            for r, blk in d_txt_blocks:
                full[r] = blk
            dist.all_reduce(full, op=dist.ReduceOp.SUM)
            d_txt = full[rank]
        g = dloss.to(F32)
        dt = ctx.in_dtypes
        return (None, (d_img * g).to(dt[0]), (d_txt * g).to(dt[1]), (sacc[0] * g).to(dt[2]), (sacc[1] * g).to(dt[3]))


class NativeSigLipLoss(nn.Module):
    """Reference ctor/forward signature: loss.py:324-338,406."""

    def __init__(self, cache_labels: bool = False, rank: int = 0, world_size: int = 1, dist_impl: Optional[str] = None,
                 chunk_size: int = 0):
        super().__init__()
        self.cache_labels, self.rank, self.world_size = cache_labels, rank, world_size
        self.dist_impl = dist_impl or "bidir"
        assert self.dist_impl in ("bidir", "shift", "reduce", "gather")
        self.chunk_size = chunk_size  # the fused kernel never materialises logits, so chunking is moot
        self._ex: Optional[comm.PeerFeatureExchange] = None

    _exchange = NativeClipLoss._exchange

    def forward(self, image_features, text_features, logit_scale, logit_bias, output_dict: bool = False):
        loss = _SigLipLossFn.apply(self, image_features, text_features, logit_scale, logit_bias)
        return {"contrastive_loss": loss} if output_dict else loss
