"""Forward / backward schedules of the two CLIP towers as sequences of libclipn kernel launches.

This is the host-side mirror of (reference) VisionTransformer.forward transformer.py:917-928,
CLIP._encode_text model.py:396-411 and ResidualAttentionBlock.forward transformer.py:319-330 — plus the
hand-scheduled backward the reference gets from autograd.  Saved per block (bf16): block input, qkv,
attention output, mid residual, c_fc pre-activation (10*d per token) + LN statistics + softmax LSE;
LayerNorm outputs and GELU outputs are recomputed in the backward (SURVEY §7 hard part 1).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops

BF16, F32 = torch.bfloat16, torch.float32


@dataclass
class TowerCfg:
    width: int
    layers: int
    heads: int
    seq: int
    causal: bool
    prefix: str          # "visual.transformer" or "transformer"
    embed_dim: int
    # vision only
    image_size: int = 0
    patch: int = 0
    # text only
    vocab: int = 0


class Scratch:
    """Named scratch tensors reused across layers / steps (caller-owned workspaces of the C ABI)."""

    def __init__(self):
        self._bufs: Dict[str, torch.Tensor] = {}

    def get(self, name: str, shape, dtype, device) -> torch.Tensor:
        t = self._bufs.get(name)
        n = 1
        for s in shape:
            n *= s
        if t is None or t.numel() < n or t.dtype != dtype or t.device != device:
            t = torch.empty(n, dtype=dtype, device=device)
            self._bufs[name] = t
        return t[:n].view(*shape)

    def clear(self):
        self._bufs.clear()


def _fast_mode() -> bool:
    """Activation policy. 'fast' (default) additionally keeps the two LayerNorm outputs and gelu(h) of every block
    (16*d bf16 per token instead of 10*d: ~122 GB instead of ~76 GB at ViT-B-32 / batch 4096), which removes 48
    LayerNorm recomputes per step and halves the HBM writes of the write-bound GELU-backward GEMM.
    CLIPN_ACT_MODE=lean restores the recomputing schedule."""
    import os
    return os.environ.get("CLIPN_ACT_MODE", "fast") != "lean"


@dataclass
class BlockSaved:
    x_in: torch.Tensor
    qkv: torch.Tensor
    att: torch.Tensor
    x_mid: torch.Tensor
    h_pre: torch.Tensor
    lse: torch.Tensor
    ln1_mean: torch.Tensor
    ln1_rstd: torch.Tensor
    ln2_mean: torch.Tensor
    ln2_rstd: torch.Tensor
    h1: Optional[torch.Tensor] = None   # ln_1 output   (fast mode)
    h2: Optional[torch.Tensor] = None   # ln_2 output   (fast mode)
    g: Optional[torch.Tensor] = None    # gelu(h_pre)   (fast mode)


@dataclass
class TowerSaved:
    batch: int = 0
    # per block: BlockSaved, or (grad checkpointing, transformer.py:397-402) just the block's input tensor
    blocks: List[object] = field(default_factory=list)
    extra: Dict[str, torch.Tensor] = field(default_factory=dict)


def _run_blocks(P, cfg: TowerCfg, x: torch.Tensor, B: int, ws: Scratch, saved: Optional[TowerSaved], checkpoint: bool):
    """The residual stack. With `checkpoint` only each block's input (d bf16 per token) is kept and the block is
    re-run in the backward — the reference's `checkpoint(r, x)` loop (transformer.py:397-402)."""
    for i in range(cfg.layers):
        pre = f"{cfg.prefix}.resblocks.{i}"
        if saved is not None and checkpoint:
            saved.blocks.append(x)
            x, _ = block_forward(P, pre, cfg, x, B, ws, False)
        else:
            x, bs = block_forward(P, pre, cfg, x, B, ws, saved is not None)
            if saved is not None:
                saved.blocks.append(bs)
    return x


def _run_blocks_backward(P, G, cfg: TowerCfg, saved: TowerSaved, dx: torch.Tensor, B: int, ws: Scratch):
    on_block = saved.extra.get("on_block_grads_ready")  # model.py: per-block gradient all-reduce (native grad sync)
    for i in reversed(range(cfg.layers)):
        pre = f"{cfg.prefix}.resblocks.{i}"
        s = saved.blocks[i]
        if isinstance(s, torch.Tensor):  # checkpointed: recompute this block's activations from its input
            _, s = block_forward(P, pre, cfg, s, B, ws, True)
        dx = block_backward(P, G, pre, cfg, s, dx, B, ws)
        saved.blocks[i] = None  # free activations as we go
        del s
        if on_block is not None:
            on_block(pre)  # every gradient of this block is final: its all-reduce overlaps the next block's backward
    return dx


# --------------------------------------------------------------------------------------------------
# residual attention block (transformer.py:319-330)
# --------------------------------------------------------------------------------------------------
def block_forward(P: Dict[str, torch.Tensor], pre: str, cfg: TowerCfg, x: torch.Tensor, batch: int, ws: Scratch,
                  save: bool) -> (torch.Tensor, Optional[BlockSaved]):
    M, d = x.shape
    dev = x.device
    keep = save and _fast_mode()
    # ln_1 -> QKV (3 F.linear on in_proj chunks == one GEMM with N = 3d, transformer.py:195-197)
    h1 = torch.empty((M, d), dtype=BF16, device=dev) if keep else ws.get("h", (M, d), BF16, dev)
    _, m1, r1 = ops.layernorm_fwd(x, P[pre + ".ln_1.weight"], P[pre + ".ln_1.bias"], out=h1, save_stats=save)
    qkv = torch.empty((M, 3 * d), dtype=BF16, device=dev) if save else ws.get("qkv", (M, 3 * d), BF16, dev)
    ops.gemm(h1, P[pre + ".attn.in_proj_weight"], bias=P[pre + ".attn.in_proj_bias"], out=qkv)
    att = torch.empty((M, d), dtype=BF16, device=dev) if save else ws.get("att", (M, d), BF16, dev)
    _, lse = ops.attention_fwd(qkv, batch, cfg.seq, cfg.heads, cfg.causal, out=att)
    # out_proj + residual (transformer.py:246,328)
    x_mid = torch.empty((M, d), dtype=BF16, device=dev)
    ops.gemm(att, P[pre + ".attn.out_proj.weight"], bias=P[pre + ".attn.out_proj.bias"], aux=x,
             epilogue=L.EPI_BIAS_RESID, out=x_mid)
    # ln_2 -> c_fc + GELU -> c_proj + residual (transformer.py:295-299,329)
    h2 = torch.empty((M, d), dtype=BF16, device=dev) if keep else ws.get("h", (M, d), BF16, dev)
    _, m2, r2 = ops.layernorm_fwd(x_mid, P[pre + ".ln_2.weight"], P[pre + ".ln_2.bias"], out=h2, save_stats=save)
    hid = P[pre + ".mlp.c_fc.weight"].shape[0]  # int(d * mlp_ratio), model.py sizes c_fc / c_proj with it
    h_pre = torch.empty((M, hid), dtype=BF16, device=dev) if save else ws.get("h_pre", (M, hid), BF16, dev)
    g = torch.empty((M, hid), dtype=BF16, device=dev) if keep else ws.get("g", (M, hid), BF16, dev)
    # fast mode keeps gelu'(h) in the `h_pre` slot (and gelu(h)), so the backward epilogue is a plain multiply;
    # lean mode keeps the pre-activation h and re-derives both in the backward (CLIPN_EPI_DGELU)
    ops.gemm(h2, P[pre + ".mlp.c_fc.weight"], bias=P[pre + ".mlp.c_fc.bias"],
             epilogue=L.EPI_BIAS_GELU_GRAD if keep else L.EPI_BIAS_GELU, out=h_pre, out2=g)
    x_out = torch.empty((M, d), dtype=BF16, device=dev)
    ops.gemm(g, P[pre + ".mlp.c_proj.weight"], bias=P[pre + ".mlp.c_proj.bias"], aux=x_mid, epilogue=L.EPI_BIAS_RESID,
             out=x_out)
    saved = BlockSaved(x, qkv, att, x_mid, h_pre, lse, m1, r1, m2, r2, h1 if keep else None, h2 if keep else None,
                       g if keep else None) if save else None
    return x_out, saved


def _wgrad(dy: torch.Tensor, x: torch.Tensor, gw: torch.Tensor):
    """gw[N_out, K_in] (fp32, +=) = dy[M, N_out]^T @ x[M, K_in]  — both operands MN-major, split-K."""
    n_out, k_in = gw.shape
    ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw,
             splits=ops.wgrad_splits(n_out, k_in, dy.shape[0]))


def block_backward(P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor], pre: str, cfg: TowerCfg, s: BlockSaved,
                   dx_out: torch.Tensor, batch: int, ws: Scratch) -> torch.Tensor:
    M, d = dx_out.shape
    dev = dx_out.device
    # ---- MLP
    hid = P[pre + ".mlp.c_fc.weight"].shape[0]
    dh = ws.get("dh", (M, hid), BF16, dev)
    # dgrad of c_proj fused with GELU backward; re-materialises g = gelu(h_pre) unless the forward kept it
    # (the c_fc bias gradient = column sums of dh is accumulated by the same epilogue)
    if s.g is not None:
        g = s.g
        ops.gemm(dx_out, P[pre + ".mlp.c_proj.weight"], b_mn=True, epilogue=L.EPI_MUL_AUX, aux=s.h_pre, out=dh,
                 col_sum=G[pre + ".mlp.c_fc.bias"])  # s.h_pre holds gelu'(h) in fast mode
    else:
        g = ws.get("g", (M, hid), BF16, dev)
        ops.gemm(dx_out, P[pre + ".mlp.c_proj.weight"], b_mn=True, epilogue=L.EPI_DGELU, aux=s.h_pre, out=dh, out2=g,
                 col_sum=G[pre + ".mlp.c_fc.bias"])
    _wgrad(dx_out, g, G[pre + ".mlp.c_proj.weight"])
    if s.h2 is not None:
        h2 = s.h2
    else:
        h2 = ws.get("h", (M, d), BF16, dev)
        ops.layernorm_fwd(s.x_mid, P[pre + ".ln_2.weight"], P[pre + ".ln_2.bias"], out=h2, save_stats=False)
    _wgrad(dh, h2, G[pre + ".mlp.c_fc.weight"])
    dh2 = ws.get("dh_small", (M, d), BF16, dev)
    ops.gemm(dh, P[pre + ".mlp.c_fc.weight"], b_mn=True, out=dh2)
    dx_mid = ws.get("dx_mid", (M, d), BF16, dev)
    # ln_2 backward streams dx_out as the residual gradient: the c_proj bias gradient (its column sums) rides along
    ops.layernorm_bwd(dh2, s.x_mid, s.ln2_mean, s.ln2_rstd, P[pre + ".ln_2.weight"], G[pre + ".ln_2.weight"],
                      G[pre + ".ln_2.bias"], resid=dx_out, out=dx_mid, resid_sum=G[pre + ".mlp.c_proj.bias"])
    # ---- attention
    datt = ws.get("dh_small", (M, d), BF16, dev)
    ops.gemm(dx_mid, P[pre + ".attn.out_proj.weight"], b_mn=True, out=datt)
    _wgrad(dx_mid, s.att, G[pre + ".attn.out_proj.weight"])
    dqkv = ws.get("dqkv", (M, 3 * d), BF16, dev)
    ops.attention_bwd(s.qkv, s.att, datt, s.lse, batch, cfg.seq, cfg.heads, cfg.causal, out=dqkv,
                      dbias=G[pre + ".attn.in_proj_bias"])
    if s.h1 is not None:
        h1 = s.h1
    else:
        h1 = ws.get("h", (M, d), BF16, dev)
        ops.layernorm_fwd(s.x_in, P[pre + ".ln_1.weight"], P[pre + ".ln_1.bias"], out=h1, save_stats=False)
    _wgrad(dqkv, h1, G[pre + ".attn.in_proj_weight"])
    dh1 = ws.get("dh_small", (M, d), BF16, dev)
    ops.gemm(dqkv, P[pre + ".attn.in_proj_weight"], b_mn=True, out=dh1)
    dx_in = torch.empty((M, d), dtype=BF16, device=dev)
    # ln_1 backward streams dx_mid: the out_proj bias gradient rides along
    ops.layernorm_bwd(dh1, s.x_in, s.ln1_mean, s.ln1_rstd, P[pre + ".ln_1.weight"], G[pre + ".ln_1.weight"],
                      G[pre + ".ln_1.bias"], resid=dx_mid, out=dx_in, resid_sum=G[pre + ".attn.out_proj.bias"])
    return dx_in


# --------------------------------------------------------------------------------------------------
# head: pool -> LN -> projection -> (optional) L2 normalize
# --------------------------------------------------------------------------------------------------
def _head_forward(x, idx, ln_w, ln_b, proj, batch, seq, normalize, saved: Optional[TowerSaved]):
    pooled_in = ops.gather_rows(x, idx, batch, seq)
    lnp, mean, rstd = ops.layernorm_fwd(pooled_in, ln_w, ln_b, save_stats=saved is not None)
    pooled = ops.gemm(lnp, proj, b_mn=True)  # `pooled @ proj` (transformer.py:923 / model.py:409)
    if normalize:
        feat, inv = ops.l2norm_fwd(pooled)
    else:
        feat, inv = pooled, None
    if saved is not None:
        saved.extra.update(pooled_in=pooled_in, lnp=lnp, head_mean=mean, head_rstd=rstd, feat=feat, inv=inv, idx=idx)
    return feat


def _head_backward(dfeat, saved: TowerSaved, ln_w, proj, g_ln_w, g_ln_b, g_proj, batch, seq, ws: Scratch, d: int):
    e = saved.extra
    if e["inv"] is not None:
        dpooled = ops.l2norm_bwd(dfeat.contiguous(), e["feat"], e["inv"])
    else:
        dpooled = dfeat.to(BF16).contiguous()
    # dproj[d, E] += lnp^T @ dpooled
    ops.gemm(e["lnp"], dpooled, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=g_proj)
    dlnp = ops.gemm(dpooled, proj)  # [B, d] = dpooled @ proj^T (proj stored [d, E] == [N, K])
    dpin = ops.layernorm_bwd(dlnp, e["pooled_in"], e["head_mean"], e["head_rstd"], ln_w, g_ln_w, g_ln_b)
    dx = torch.empty((batch * seq, d), dtype=BF16, device=dfeat.device)
    ops.scatter_rows(dpin, e["idx"], batch, seq, out=dx)
    return dx


# --------------------------------------------------------------------------------------------------
# vision tower
# --------------------------------------------------------------------------------------------------
def _conv1_weight_rows(w4: torch.Tensor, patch: int) -> torch.Tensor:
    """conv1.weight [d,3,P,P] as the GEMM's [d, K] operand; for patch 14 a zero-padded copy with K = 592 (the same
    im2row kernel with one 'patch' per output channel does the padding)."""
    if patch % 8 == 0:
        return w4.view(w4.shape[0], -1)
    return ops.patchify(w4, patch)


def vision_forward(P, cfg: TowerCfg, image: torch.Tensor, normalize: bool, ws: Scratch, save: bool,
                   checkpoint: bool = False):
    B = image.shape[0]
    if image.dtype != BF16:
        image = image.to(BF16)  # prepare_batch casts inputs to the model's input dtype (base_task.py:148-152)
    image = image.contiguous()
    grid = cfg.image_size // cfg.patch
    npatch = grid * grid
    d = cfg.width
    saved = TowerSaved(batch=B) if save else None
    patches = ops.patchify(image, cfg.patch)
    w = _conv1_weight_rows(P["visual.conv1.weight"], cfg.patch)
    pe = ops.gemm(patches, w)
    x0 = ops.vision_embed_fwd(pe, P["visual.class_embedding"], P["visual.positional_embedding"], B, npatch)
    x, m0, r0 = ops.layernorm_fwd(x0, P["visual.ln_pre.weight"], P["visual.ln_pre.bias"], save_stats=save)
    if save:
        saved.extra.update(patches=patches, x0=x0, pre_mean=m0, pre_rstd=r0)
    x = _run_blocks(P, cfg, x, B, ws, saved, checkpoint)
    feat = _head_forward(x, None, P["visual.ln_post.weight"], P["visual.ln_post.bias"], P["visual.proj"], B, cfg.seq,
                         normalize, saved)
    return feat, saved


def vision_backward(P, G, cfg: TowerCfg, saved: TowerSaved, dfeat: torch.Tensor, ws: Scratch):
    B = saved.batch
    d = cfg.width
    grid = cfg.image_size // cfg.patch
    npatch = grid * grid
    dx = _head_backward(dfeat, saved, P["visual.ln_post.weight"], P["visual.proj"], G["visual.ln_post.weight"],
                        G["visual.ln_post.bias"], G["visual.proj"], B, cfg.seq, ws, d)
    dx = _run_blocks_backward(P, G, cfg, saved, dx, B, ws)
    e = saved.extra
    dx0 = ops.layernorm_bwd(dx, e["x0"], e["pre_mean"], e["pre_rstd"], P["visual.ln_pre.weight"],
                            G["visual.ln_pre.weight"], G["visual.ln_pre.bias"])
    dpe = ops.vision_embed_bwd(dx0, G["visual.class_embedding"], G["visual.positional_embedding"], B, npatch)
    gw = G["visual.conv1.weight"].view(d, -1)
    k, kp = gw.shape[1], e["patches"].shape[1]
    if kp == k:
        _wgrad(dpe, e["patches"], gw)
    else:  # patch 14: the im2row matrix is K-padded; accumulate into a padded scratch and fold the valid columns back
        gw_pad = ws.get("conv1_grad_pad", (d, kp), F32, dpe.device)
        gw_pad.zero_()
        _wgrad(dpe, e["patches"], gw_pad)
        ops.accum_rows_f32(gw, gw_pad, k)


# --------------------------------------------------------------------------------------------------
# text tower
# --------------------------------------------------------------------------------------------------
def text_forward(P, cfg: TowerCfg, text: torch.Tensor, normalize: bool, ws: Scratch, save: bool,
                 checkpoint: bool = False):
    B = text.shape[0]
    text = text.contiguous()
    saved = TowerSaved(batch=B) if save else None
    x, eot = ops.text_embed_fwd(text, P["token_embedding.weight"], P["positional_embedding"])
    if save:
        saved.extra.update(text=text)
    x = _run_blocks(P, cfg, x, B, ws, saved, checkpoint)
    feat = _head_forward(x, eot, P["ln_final.weight"], P["ln_final.bias"], P["text_projection"], B, cfg.seq, normalize,
                         saved)
    return feat, saved


def text_backward(P, G, cfg: TowerCfg, saved: TowerSaved, dfeat: torch.Tensor, ws: Scratch):
    B = saved.batch
    d = cfg.width
    dx = _head_backward(dfeat, saved, P["ln_final.weight"], P["text_projection"], G["ln_final.weight"],
                        G["ln_final.bias"], G["text_projection"], B, cfg.seq, ws, d)
    dx = _run_blocks_backward(P, G, cfg, saved, dx, B, ws)
    ops.text_embed_bwd(saved.extra["text"], dx, G["token_embedding.weight"], G["positional_embedding"])
