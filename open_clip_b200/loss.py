"""NativeClipLoss / NativeSigLipLoss — drop-ins for the reference ClipLoss (loss.py:57-141) and SigLipLoss
(loss.py:314-489): same constructor kwargs, same forward signature, same returned value / dict, same gradient
conventions for every (local_loss, gather_with_grad) combination — computed by libclipn's fused kernels.

Forward: the [B x N] logits are never materialised.  ONE launch of the peer-streaming tcgen05 kernel produces, for
both directions, the online log-sum-exp (ClipLoss) or the softplus sum (SigLipLoss) of this rank's rows against
EVERY rank's features, reading each peer's buffer tile by tile over NVLink (comm.FeatureGather) — the all-gathers of
`gather_features` are fused into the GEMM — and leaves a local copy of the gathered operands behind.
Backward: d(logits) tiles are recomputed from the LSE vectors against the local gathered copy (the only cross-rank
exchange is 2*N fp32 for ClipLoss, nothing for SigLipLoss), written once in bf16 and contracted by split-K tcgen05
GEMMs.  This reproduces the reduce-scatter semantics of `gather_with_grad` exactly (each rank's feature gradient =
sum over ALL ranks' losses) without moving any [N,E] gradient across NVLink.

Supported envelope: any world size / batch; shapes or transports the fused kernel does not take (world > 8, several
nodes, per-rank batch not a multiple of 128, embed dim not a multiple of 64) run the same math after an NCCL
all-gather (comm.FeatureGather, mode "nccl").  Hard limits, raised as ClipnError: CUDA tensors only, embed dim % 8 == 0
(any batch size: ragged d(logits) rows are padded to a 16-byte pitch internally).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import comm, ops
from ._lib import ClipnError

BF16, F32 = torch.bfloat16, torch.float32


def _check_inputs(name: str, image_features, text_features, world: int):
    if not image_features.is_cuda:
        raise ClipnError(f"{name} runs on CUDA tensors only; there is no CPU fallback")
    if image_features.dim() != 2 or image_features.shape != text_features.shape:
        raise ClipnError(f"{name}: image/text features must both be [B, E], got {tuple(image_features.shape)} and "
                         f"{tuple(text_features.shape)}")
    b, e = image_features.shape
    if e % 8 != 0:
        raise ClipnError(f"{name}: the embed dim ({e}) must be a multiple of 8 (16-byte rows for the TMA tensor maps)")


def _gathered(module, img, txt, B, E, W):
    """Shared forward plumbing: returns (img_loc, txt_loc, all_img, all_txt, fused, ptrs) where `fused` says whether
    the peer-streaming kernel runs (ptrs = (img_ptrs, txt_ptrs, gather_img, gather_txt)) or the operands were gathered
    up front (W == 1: the local tensors themselves)."""
    if W == 1:
        fused = ops.peer_gemm_tile_n(1, B, E) != 0
        return img, txt, img, txt, fused, ([img.data_ptr()], [txt.data_ptr()], None, None)
    g = module._gather(B, E, img.device)
    if g.mode == "peer":
        # the ROW operands stay the ordinary local tensors: the symmetric buffer is only what the peers read (measured
        # on 2 x B200: streaming the row operand out of the symmetric mapping made the GEMM 2.4x slower)
        _, _, img_ptrs, txt_ptrs = g.publish(img, txt)
        return img, txt, g.all_img, g.all_txt, True, (img_ptrs, txt_ptrs, g.all_img, g.all_txt)
    g.gather_nccl(img, txt)
    return img, txt, g.all_img, g.all_txt, False, None


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "NativeClipLoss", image_features, text_features, logit_scale, logit_bias):
        # loss.py:100-116 adds logit_bias to every logit; softmax cross-entropy is invariant to a constant added to a
        # whole row, so the value is unchanged and d loss / d logit_bias = sum(softmax - onehot) = 0.
        ctx.bias_like = logit_bias.detach() if logit_bias is not None else None
        W, rank = module.world_size, module.rank
        _check_inputs("NativeClipLoss", image_features, text_features, W)
        B, E = image_features.shape
        img = image_features.detach().to(BF16).contiguous()
        txt = text_features.detach().to(BF16).contiguous()
        scale = logit_scale.detach().to(F32).reshape(1).contiguous()  # stays on the device: no host sync
        img, txt, all_img, all_txt, fused, ptrs = _gathered(module, img, txt, B, E, W)
        off = rank * B if W > 1 else 0
        if fused:
            img_ptrs, txt_ptrs, g_img, g_txt = ptrs
            lse, loss1 = ops.clip_fwd_fused(img, txt, txt_ptrs, img_ptrs, rank, scale, g_txt, g_img)
            lse_i, lse_t, local = lse[0], lse[1], loss1[0]
        else:
            lse_i, pos_i = ops.clip_lse_fwd(img, all_txt, scale, off)  # logits_per_image rows (loss.py:103)
            lse_t, pos_t = ops.clip_lse_fwd(txt, all_img, scale, off)  # logits_per_text rows  (loss.py:104)
            local = ((lse_i - pos_i).mean() + (lse_t - pos_t).mean()) * 0.5
        gscale, col_w, global_value = comm.clip_grad_convention(module.local_loss, module.gather_with_grad, B, W)
        loss = local
        if global_value:
            loss = local.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            loss = loss / W
        ctx.meta = (B, E, W, off, scale, gscale, col_w, global_value)
        # the backward reads the local gathered copies; they are reused by the next forward of this module, so a
        # second forward before this backward is detected through the generation counter
        module._generation += 1
        ctx.module, ctx.generation = module, module._generation
        ctx.saved = (img, txt, all_img, all_txt, lse_i, lse_t)
        ctx.in_dtypes = (image_features.dtype, text_features.dtype, logit_scale.dtype)
        return loss.to(image_features.dtype) if image_features.dtype != F32 else loss

    @staticmethod
    def backward(ctx, dloss):
        B, E, W, off, scale, gscale, col_w, global_value = ctx.meta
        img, txt, all_img, all_txt, lse_i, lse_t = ctx.saved
        if W > 1 and ctx.generation != ctx.module._generation:
            raise ClipnError("NativeClipLoss.backward: the module ran another multi-rank forward since this loss was "
                             "computed; its gathered feature buffers were overwritten (use one loss module per "
                             "concurrent graph)")
        if W > 1 and col_w != 0.0:
            both = comm.all_gather_vectors(torch.stack([lse_i, lse_t]))  # [2, N]
            all_lse_i, all_lse_t = both[0], both[1]
        else:
            all_lse_i, all_lse_t = lse_i, lse_t
        acc = torch.zeros(4, dtype=F32, device=img.device)
        # d logits_per_image[B,N] (rows = my images): row softmax uses my image LSE, column term uses every text's LSE
        # The bf16 d(logits) tiles hold the softmax parts centred on their mean (1+w)/N and no one-hot; both enter the
        # fp32 accumulator directly: (1+w) * gscale * scale * (mean of the gathered operand - its row off+m, which is
        # this rank's own OTHER-modality feature m).  bf16 never rounds a value ~2 or ~1/N, only deviations.
        N = W * B
        coef = scale * ((1.0 + col_w) * gscale)  # [1] fp32 device scalar
        mean_t = ops.colsum(all_txt, torch.zeros(E, dtype=F32, device=img.device)) / N
        mean_i = ops.colsum(all_img, torch.zeros(E, dtype=F32, device=img.device)) / N
        # per-row centre of the d logit_scale sum (clipn.h): the positive pair's cosine, the same for both directions
        centre = (img.float() * txt.float()).sum(dim=-1).contiguous()
        dl_i = ops.clip_dlogits(img, all_txt, scale, off, lse_i, all_lse_t if col_w else None, col_w, gscale, acc[0:2],
                                row_centre=centre)
        d_img = ops.clip_dfeat(dl_i, all_txt, scale, n=N, init=(mean_t - txt.float()) * coef)
        del dl_i
        dl_t = ops.clip_dlogits(txt, all_img, scale, off, lse_t, all_lse_i if col_w else None, col_w, gscale, acc[2:4],
                                row_centre=centre)
        d_txt = ops.clip_dfeat(dl_t, all_img, scale, n=N, init=(mean_i - img.float()) * coef)
        del dl_t
        # d loss / d logit_scale = sum_{dir} sum (P_row - onehot) * <row, col> / (2B)   (own loss only)
        d_scale = (acc[0] + acc[2]) * ((1.0 / (2 * B)) / gscale)
        if global_value:
            d_scale = d_scale.clone()
            dist.all_reduce(d_scale, op=dist.ReduceOp.SUM)
            d_scale = d_scale / W
        g = dloss.to(F32)
        d_img = (d_img * g).to(ctx.in_dtypes[0])
        d_txt = (d_txt * g).to(ctx.in_dtypes[1])
        d_bias = torch.zeros_like(ctx.bias_like) if ctx.bias_like is not None else None
        return None, d_img, d_txt, (d_scale * g).to(ctx.in_dtypes[2]), d_bias


class _LossBase(nn.Module):
    def _init_exchange(self):
        self._fg: Optional[comm.FeatureGather] = None
        self._generation = 0

    def _gather(self, batch: int, embed: int, device) -> comm.FeatureGather:
        if self._fg is None or self._fg.batch != batch or self._fg.embed != embed:
            self._fg = comm.FeatureGather(batch, embed, device)
        return self._fg

    @property
    def exchange_mode(self) -> str:
        """'peer' (fused NVLink reads), 'nccl' (fallback) or 'local' (world size 1 / not yet run)."""
        return self._fg.mode if self._fg is not None else "local"


class NativeClipLoss(_LossBase):
    """Reference ctor/forward signature: loss.py:59-72,118-141."""

    def __init__(self, local_loss: bool = False, gather_with_grad: bool = False, cache_labels: bool = False,
                 rank: int = 0, world_size: int = 1):
        super().__init__()
        self.local_loss, self.gather_with_grad, self.cache_labels = local_loss, gather_with_grad, cache_labels
        self.rank, self.world_size = rank, world_size
        self._init_exchange()

    def forward(self, image_features, text_features, logit_scale, logit_bias=None, output_dict: bool = False):
        loss = _ClipLossFn.apply(self, image_features, text_features, logit_scale, logit_bias)
        return {"contrastive_loss": loss} if output_dict else loss


class _SigLipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "NativeSigLipLoss", need_grad: bool, image_features, text_features, logit_scale,
                logit_bias):
        W, rank = module.world_size, module.rank
        _check_inputs("NativeSigLipLoss", image_features, text_features, W)
        B, E = image_features.shape
        img = image_features.detach().to(BF16).contiguous()
        txt = text_features.detach().to(BF16).contiguous()
        scale = logit_scale.detach().to(F32).reshape(1).contiguous()
        bias = logit_bias.detach().to(F32).reshape(1).contiguous()
        loss_acc = torch.zeros(1, dtype=F32, device=img.device)
        sacc = torch.zeros(2, dtype=F32, device=img.device)
        gscale = 1.0 / B  # loss.py:366 `.sum() / image_features.shape[0]`
        img, txt, all_img, all_txt, fused, ptrs = _gathered(module, img, txt, B, E, W)
        off = rank * B if W > 1 else 0
        N = W * B
        # Every other rank's text block is visited exactly once as a negative-only block (loss.py:410-487: all four
        # dist_impl's are the same sum) = this rank's image rows against ALL text columns with one positive per row.
        # The text gradient needs no reverse exchange: d loss_total / d txt_r[j] = sum over ALL image rows i of
        # dz(i, j) * img_i, and dz depends only on the pair -> this rank's text rows against all image columns.
        if fused:
            img_ptrs, txt_ptrs, g_img, g_txt = ptrs
            dl_i, dl_t = ops.siglip_fwd_fused(img, txt, txt_ptrs, img_ptrs, rank, scale, bias, gscale, g_txt, g_img,
                                              loss_acc, sacc, need_grad)
        else:
            dl_i = ops.siglip_dir(img, all_txt, scale, bias, off, gscale, loss_acc, sacc, need_grad)
            dl_t = ops.siglip_dir(txt, all_img, scale, bias, off, gscale, None, None, True) if need_grad else None
        d_img = d_txt = None
        if need_grad:
            d_img = ops.clip_dfeat(dl_i, all_txt, scale, n=N)  # scale * dl_i @ all_txt
            d_txt = ops.clip_dfeat(dl_t, all_img, scale, n=N)  # scale * dl_t @ all_img
        ctx.saved = (d_img, d_txt, sacc)
        ctx.in_dtypes = tuple(t.dtype for t in (image_features, text_features, logit_scale, logit_bias))
        loss = loss_acc[0]
        return loss.to(image_features.dtype) if image_features.dtype != F32 else loss

    @staticmethod
    def backward(ctx, dloss):
        d_img, d_txt, sacc = ctx.saved
        if d_img is None:
            raise ClipnError("NativeSigLipLoss.backward: the forward ran without gradients enabled")
        g = dloss.to(F32)
        dt = ctx.in_dtypes
        return (None, None, (d_img * g).to(dt[0]), (d_txt * g).to(dt[1]), (sacc[0] * g).to(dt[2]),
                (sacc[1] * g).to(dt[3]))


class NativeSigLipLoss(_LossBase):
    """Reference ctor/forward signature: loss.py:324-338,406."""

    def __init__(self, cache_labels: bool = False, rank: int = 0, world_size: int = 1, dist_impl: Optional[str] = None,
                 chunk_size: int = 0):
        super().__init__()
        self.cache_labels, self.rank, self.world_size = cache_labels, rank, world_size
        self.dist_impl = dist_impl or "bidir"
        assert self.dist_impl in ("bidir", "shift", "reduce", "gather")
        self.chunk_size = chunk_size  # the fused kernel never materialises logits, so chunking is moot
        self._init_exchange()

    def forward(self, image_features, text_features, logit_scale, logit_bias, output_dict: bool = False):
        # the gradient GEMMs run inside the forward (d(logits) is produced by the same epilogue as the value), so
        # whether they are needed is decided here, where grad mode is still visible
        need_grad = torch.is_grad_enabled() and any(
            t.requires_grad for t in (image_features, text_features, logit_scale, logit_bias))
        loss = _SigLipLossFn.apply(self, need_grad, image_features, text_features, logit_scale, logit_bias)
        return {"contrastive_loss": loss} if output_dict else loss
