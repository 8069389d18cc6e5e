"""CPU oracle for the CLIP train-step hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain-PyTorch (CPU, fp32 or emulated-bf16) restatement of the reference's
algorithm for the north-star path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``open_clip_b200`` never does (it fails loudly when the CUDA library is missing).

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference from
``/root/reference/src`` in the authoring container, checks this restatement against it
(forward, loss, every parameter gradient, 1- and multi-rank losses via gloo) and writes
the fixtures under ``tests/golden/`` that the GPU-box tests compare against.  The
reference's own tests pin only the SigLIP chunked==unchunked identity
(``tests/test_siglip_chunked_loss.py``), which ``tests/test_oracle.py`` re-checks here.

Every function cites the reference lines it follows (paths relative to
``/root/reference/src/open_clip``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# configuration (mirrors model_configs/*.json; model.py:37-150 CLIPVisionCfg/CLIPTextCfg)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ClipCfg:
    embed_dim: int = 512
    image_size: int = 224
    patch_size: int = 32
    v_width: int = 768
    v_layers: int = 12
    v_head_width: int = 64
    t_ctx: int = 77
    t_vocab: int = 49408
    t_width: int = 512
    t_heads: int = 8
    t_layers: int = 12
    mlp_ratio: float = 4.0

    @property
    def v_heads(self) -> int:  # model.py:209 heads = width // head_width
        return self.v_width // self.v_head_width

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def v_tokens(self) -> int:
        return self.grid * self.grid + 1


CONFIGS: Dict[str, ClipCfg] = {
    # model_configs/ViT-B-32.json
    "ViT-B-32": ClipCfg(),
    # model_configs/ViT-B-16.json
    "ViT-B-16": ClipCfg(patch_size=16),
    # model_configs/ViT-L-14-336.json
    "ViT-L-14-336": ClipCfg(embed_dim=768, image_size=336, patch_size=14, v_width=1024, v_layers=24,
                            t_width=768, t_heads=12, t_layers=12),
    # ViT-L-14-336 geometry (patch 14, 577 tokens, width 1024/768) at depth 2+2: reference autograd on CPU in seconds
    "ViT-L-14-336-d2": ClipCfg(embed_dim=768, image_size=336, patch_size=14, v_width=1024, v_layers=2,
                               t_width=768, t_heads=12, t_layers=2),
    # small shape-compatible config for fast fixtures (head_dim stays 64 like every native ViT)
    "tiny": ClipCfg(embed_dim=128, image_size=64, patch_size=16, v_width=128, v_layers=2,
                    t_ctx=20, t_vocab=512, t_width=128, t_heads=2, t_layers=2),
}


# --------------------------------------------------------------------------------------
# parameter construction (names/shapes: SURVEY §8b; init: transformer.py:145-155,1664-1685,
# 641-645,719-720; model.py:326)
# --------------------------------------------------------------------------------------
def param_shapes(cfg: ClipCfg) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    s["positional_embedding"] = (cfg.t_ctx, cfg.t_width)
    s["text_projection"] = (cfg.t_width, cfg.embed_dim)
    s["logit_scale"] = ()
    s["visual.class_embedding"] = (cfg.v_width,)
    s["visual.positional_embedding"] = (cfg.v_tokens, cfg.v_width)
    s["visual.proj"] = (cfg.v_width, cfg.embed_dim)
    s["visual.conv1.weight"] = (cfg.v_width, 3, cfg.patch_size, cfg.patch_size)
    s["visual.ln_pre.weight"] = (cfg.v_width,)
    s["visual.ln_pre.bias"] = (cfg.v_width,)

    def blocks(prefix: str, d: int, layers: int):
        h = int(d * cfg.mlp_ratio)
        for i in range(layers):
            p = f"{prefix}.resblocks.{i}"
            s[f"{p}.ln_1.weight"] = (d,)
            s[f"{p}.ln_1.bias"] = (d,)
            s[f"{p}.attn.in_proj_weight"] = (3 * d, d)
            s[f"{p}.attn.in_proj_bias"] = (3 * d,)
            s[f"{p}.attn.out_proj.weight"] = (d, d)
            s[f"{p}.attn.out_proj.bias"] = (d,)
            s[f"{p}.ln_2.weight"] = (d,)
            s[f"{p}.ln_2.bias"] = (d,)
            s[f"{p}.mlp.c_fc.weight"] = (h, d)
            s[f"{p}.mlp.c_fc.bias"] = (h,)
            s[f"{p}.mlp.c_proj.weight"] = (d, h)
            s[f"{p}.mlp.c_proj.bias"] = (d,)

    blocks("visual.transformer", cfg.v_width, cfg.v_layers)
    s["visual.ln_post.weight"] = (cfg.v_width,)
    s["visual.ln_post.bias"] = (cfg.v_width,)
    blocks("transformer", cfg.t_width, cfg.t_layers)
    s["token_embedding.weight"] = (cfg.t_vocab, cfg.t_width)
    s["ln_final.weight"] = (cfg.t_width,)
    s["ln_final.bias"] = (cfg.t_width,)
    return s


def is_lowp_param(name: str) -> bool:
    """Which parameters `convert_weights_to_lp` (model.py:738-765) casts to bf16 under
    --precision bf16: Conv/Linear weights+biases, Attention.in_proj_*, text_projection,
    visual.proj.  LayerNorm affine, embeddings and logit_scale stay fp32."""
    if name in ("text_projection", "visual.proj", "visual.conv1.weight"):
        return True
    return (".attn." in name) or (".mlp." in name)


def init_params(cfg: ClipCfg, seed: int = 0, init_logit_scale: float = math.log(1 / 0.07),
                init_logit_bias: Optional[float] = None, bias_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """fp32 parameters drawn from the reference's init distributions with a private
    torch.Generator (so the same dict is reproducible on any box).  `bias_std>0` perturbs
    the zero-initialised biases / LN affine so parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(cfg)
    p: Dict[str, torch.Tensor] = {}

    def normal(shape, std):
        return torch.randn(shape, generator=g) * std

    def xavier_uniform(shape):  # transformer.py:148 (nn.init.xavier_uniform_ on in_proj_weight)
        fan_out, fan_in = shape
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=g) * 2 - 1) * a

    def kaiming_linear(shape):  # nn.Linear / nn.Conv2d default init (kaiming_uniform a=sqrt(5))
        fan_in = 1
        for s_ in shape[1:]:
            fan_in *= s_
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    vs = cfg.v_width ** -0.5
    for name, shape in shapes.items():
        if name == "logit_scale":
            t = torch.tensor(float(init_logit_scale))
        elif name == "positional_embedding":
            t = normal(shape, 0.01)  # transformer.py:1666
        elif name == "token_embedding.weight":
            t = normal(shape, 0.02)  # transformer.py:1665
        elif name == "text_projection":
            t = normal(shape, cfg.t_width ** -0.5)  # transformer.py:1685
        elif name in ("visual.class_embedding", "visual.positional_embedding", "visual.proj"):
            t = normal(shape, vs)  # transformer.py:641-645,719
        elif name == "visual.conv1.weight":
            t = kaiming_linear(shape)
        elif name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name.endswith("ln_pre.weight") \
                or name.endswith("ln_post.weight") or name.endswith("ln_final.weight"):
            t = torch.ones(shape) + (normal(shape, bias_std) if bias_std else 0)
        elif name.endswith(".bias"):
            if name.endswith("mlp.c_fc.bias") or name.endswith("mlp.c_proj.bias"):
                # nn.Linear default bias init U(-1/sqrt(fan_in), +1/sqrt(fan_in)) (reference keeps it)
                fan_in = shapes[name[:-4] + "weight"][1]
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
            else:
                t = torch.zeros(shape)  # in_proj_bias / out_proj.bias zero (transformer.py:153-155); LN bias 0
            if bias_std:
                t = t + normal(shape, bias_std)
        elif name.startswith("visual.") and ".attn.in_proj_weight" in name:
            t = xavier_uniform(shape)
        elif name.startswith("visual."):
            t = kaiming_linear(shape)  # vision tower keeps torch defaults (transformer.py:755-773 `pass`)
        else:  # text tower (transformer.py:1670-1677)
            d, L = cfg.t_width, cfg.t_layers
            proj_std = (d ** -0.5) * ((2 * L) ** -0.5)
            attn_std = d ** -0.5
            fc_std = (2 * d) ** -0.5
            if "in_proj_weight" in name:
                t = normal(shape, attn_std)
            elif "c_fc.weight" in name:
                t = normal(shape, fc_std)
            else:
                t = normal(shape, proj_std)
        p[name] = t.float().contiguous()
    if init_logit_bias is not None:
        p["logit_bias"] = torch.tensor(float(init_logit_bias))
    return p


def cast_params(params: Dict[str, torch.Tensor], precision: str) -> Dict[str, torch.Tensor]:
    """'fp32' -> as is; 'bf16' -> the reference's --precision bf16 dtype contract."""
    if precision == "fp32":
        return {k: v.clone() for k, v in params.items()}
    assert precision == "bf16"
    return {k: (v.to(torch.bfloat16) if is_lowp_param(k) else v.clone()) for k, v in params.items()}


# --------------------------------------------------------------------------------------
# forward (functional)
# --------------------------------------------------------------------------------------
def _ln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    # layers.py:11-26 — LayerNormFp32 (fp32 compute, cast back) when x is bf16; eps 1e-5
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)


def _attention(x: torch.Tensor, p: Dict[str, torch.Tensor], pre: str, heads: int,
               attn_mask: Optional[torch.Tensor]) -> torch.Tensor:
    # transformer.py:157-248 (cross path with k_x = v_x = x: three linears on weight chunks)
    N, L, C = x.shape
    w, b = p[f"{pre}.attn.in_proj_weight"], p[f"{pre}.attn.in_proj_bias"]
    wq, wk, wv = w.chunk(3, dim=0)
    bq, bk, bv = b.chunk(3)
    q = F.linear(x, wq, bq).reshape(N, L, heads, C // heads).transpose(1, 2)
    k = F.linear(x, wk, bk).reshape(N, L, heads, C // heads).transpose(1, 2)
    v = F.linear(x, wv, bv).reshape(N, L, heads, C // heads).transpose(1, 2)
    m = attn_mask.to(q.dtype) if attn_mask is not None else None  # transformer.py:316
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=m, scale=(C // heads) ** -0.5)  # :223-228
    o = o.transpose(1, 2).reshape(N, L, C)
    return F.linear(o, p[f"{pre}.attn.out_proj.weight"], p[f"{pre}.attn.out_proj.bias"])  # :246


def _block(x: torch.Tensor, p: Dict[str, torch.Tensor], pre: str, heads: int,
           attn_mask: Optional[torch.Tensor]) -> torch.Tensor:
    # transformer.py:319-330
    x = x + _attention(_ln(x, p[f"{pre}.ln_1.weight"], p[f"{pre}.ln_1.bias"]), p, pre, heads, attn_mask)
    h = F.linear(_ln(x, p[f"{pre}.ln_2.weight"], p[f"{pre}.ln_2.bias"]),
                 p[f"{pre}.mlp.c_fc.weight"], p[f"{pre}.mlp.c_fc.bias"])
    h = F.gelu(h)  # nn.GELU() exact erf (model.py:183)
    return x + F.linear(h, p[f"{pre}.mlp.c_proj.weight"], p[f"{pre}.mlp.c_proj.bias"])


def causal_mask(L: int) -> torch.Tensor:
    # transformer.py:1716-1722
    return torch.full((L, L), float("-inf")).triu_(1)


def encode_image(p: Dict[str, torch.Tensor], cfg: ClipCfg, image: torch.Tensor, normalize: bool = False,
                 taps: Optional[dict] = None) -> torch.Tensor:
    """VisionTransformer.forward, transformer.py:917-928 (+_embeds :793-808, _pool :810-833)."""
    dt = p["visual.conv1.weight"].dtype
    x = F.conv2d(image.to(dt), p["visual.conv1.weight"], stride=cfg.patch_size)  # :794
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # :795-796
    cls = p["visual.class_embedding"].to(dt).view(1, 1, -1).expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1)  # :799
    x = x + p["visual.positional_embedding"].to(dt)  # :801
    x = _ln(x, p["visual.ln_pre.weight"], p["visual.ln_pre.bias"])  # :807
    if taps is not None:
        taps["v_embed"] = x
    for i in range(cfg.v_layers):
        x = _block(x, p, f"visual.transformer.resblocks.{i}", cfg.v_heads, None)
        if taps is not None:
            taps[f"v_block{i}"] = x
    x = _ln(x, p["visual.ln_post.weight"], p["visual.ln_post.bias"])  # :830
    pooled = x[:, 0]  # :787 'tok'
    pooled = pooled @ p["visual.proj"]  # :923
    return F.normalize(pooled, dim=-1) if normalize else pooled  # model.py:391


def encode_text(p: Dict[str, torch.Tensor], cfg: ClipCfg, text: torch.Tensor, normalize: bool = False,
                taps: Optional[dict] = None) -> torch.Tensor:
    """CLIP._encode_text, model.py:396-411."""
    dt = p["transformer.resblocks.0.mlp.c_fc.weight"].dtype  # transformer.py:536-538 get_cast_dtype
    x = F.embedding(text, p["token_embedding.weight"]).to(dt)  # model.py:399
    x = x + p["positional_embedding"].to(dt)  # :401
    if taps is not None:
        taps["t_embed"] = x
    mask = causal_mask(cfg.t_ctx)
    for i in range(cfg.t_layers):
        x = _block(x, p, f"transformer.resblocks.{i}", cfg.t_heads, mask)
        if taps is not None:
            taps[f"t_block{i}"] = x
    x = _ln(x, p["ln_final.weight"], p["ln_final.bias"])  # :403
    pooled = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]  # transformer.py:941-944
    pooled = pooled @ p["text_projection"]  # model.py:409
    return F.normalize(pooled, dim=-1) if normalize else pooled


def clip_forward(p: Dict[str, torch.Tensor], cfg: ClipCfg, image: Optional[torch.Tensor],
                 text: Optional[torch.Tensor]) -> Dict[str, Optional[torch.Tensor]]:
    """CLIP.forward with output_dict=True, model.py:528-548."""
    out = {
        "image_features": encode_image(p, cfg, image, True) if image is not None else None,
        "text_features": encode_text(p, cfg, text, True) if text is not None else None,
        "logit_scale": p["logit_scale"].exp(),
    }
    if "logit_bias" in p:
        out["logit_bias"] = p["logit_bias"].clone()
    return out


# --------------------------------------------------------------------------------------
# losses.  Multi-rank semantics are restated WITHOUT a process group: the caller passes
# every rank's features and gets the per-rank loss values the reference produces.
# --------------------------------------------------------------------------------------
def clip_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
              logit_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ClipLoss.forward, world_size == 1 (loss.py:108-141)."""
    lpi = logit_scale * image_features @ text_features.T
    lpt = logit_scale * text_features @ image_features.T
    if logit_bias is not None:
        lpi = lpi + logit_bias
        lpt = lpt + logit_bias
    labels = torch.arange(lpi.shape[0], device=lpi.device)
    return (F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels)) / 2


def _clip_loss_one_rank(r: int, image_features, text_features, logit_scale, local_loss: bool,
                        gather_with_grad: bool) -> torch.Tensor:
    """ClipLoss value on rank r of world_size = len(image_features) (loss.py:29-54,91-141)."""
    if gather_with_grad:
        imgs, txts = list(image_features), list(text_features)
    else:
        imgs = [f.detach() for f in image_features]
        txts = [f.detach() for f in text_features]
        if not local_loss:  # loss.py:47-50 splice the grad-carrying local tensors back in
            imgs[r], txts[r] = image_features[r], text_features[r]
    all_i, all_t = torch.cat(imgs), torch.cat(txts)
    if local_loss:  # loss.py:102-104
        lpi = logit_scale * image_features[r] @ all_t.T
        lpt = logit_scale * text_features[r] @ all_i.T
        B = lpi.shape[0]
        labels = torch.arange(B, device=lpi.device) + B * r  # loss.py:82-83
    else:  # loss.py:106-107
        lpi = logit_scale * all_i @ all_t.T
        lpt = lpi.T
        labels = torch.arange(lpi.shape[0], device=lpi.device)
    return (F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels)) / 2


def clip_loss_ranks(image_features: Sequence[torch.Tensor], text_features: Sequence[torch.Tensor],
                    logit_scale: torch.Tensor, local_loss: bool, gather_with_grad: bool) -> List[torch.Tensor]:
    """Per-rank ClipLoss values for world_size = len(image_features) (loss.py:29-54,91-141).
    Gradient convention: with gather_with_grad the gathered tensors carry grad
    to every rank's features; without it only the local slice does (others are detached)."""
    return [_clip_loss_one_rank(r, image_features, text_features, logit_scale, local_loss, gather_with_grad)
            for r in range(len(image_features))]


def clip_loss_rank_grads(image_features: Sequence[torch.Tensor], text_features: Sequence[torch.Tensor],
                         logit_scale: torch.Tensor, rank: int, local_loss: bool, gather_with_grad: bool):
    """What ONE rank of the reference sees after loss.backward(): (loss_rank, d image_features[rank],
    d text_features[rank], d logit_scale).  Feature gradients are those of the SUM of every rank's loss (each rank
    back-propagates its own loss and the gather's autograd routes the pieces to their owners, loss.py:23-26,47-50);
    the logit_scale gradient is that of this rank's loss only.  Same numbers as differentiating
    sum(clip_loss_ranks(...)), but one rank's graph at a time (memory-lean: benchmark-sized batches), on whatever
    device the inputs live."""
    W = len(image_features)
    img = [f.detach() for f in image_features]
    txt = [f.detach() for f in text_features]
    img[rank] = img[rank].clone().requires_grad_(True)
    txt[rank] = txt[rank].clone().requires_grad_(True)
    d_scale = None
    value = None
    for r in range(W):
        scale = logit_scale.detach().clone().requires_grad_(True)
        loss = _clip_loss_one_rank(r, img, txt, scale, local_loss, gather_with_grad)
        if not loss.requires_grad:
            continue
        loss.backward()
        if r == rank:
            value, d_scale = loss.detach(), scale.grad.detach()
    if value is None:
        value = _clip_loss_one_rank(rank, img, txt, logit_scale.detach(), local_loss, gather_with_grad).detach()
    return value, img[rank].grad, txt[rank].grad, d_scale


def siglip_block_loss(image_features, text_features, logit_scale, logit_bias, negative_only=False):
    """SigLipLoss._loss (loss.py:351-367)."""
    logits = logit_scale * image_features @ text_features.T
    if logit_bias is not None:
        logits = logits + logit_bias
    n = image_features.shape[0]
    labels = -torch.ones((n, n), dtype=image_features.dtype, device=image_features.device)
    if not negative_only:
        labels = 2 * torch.eye(n, dtype=image_features.dtype, device=image_features.device) + labels
    return -F.logsigmoid(labels * logits).sum() / n


def _siglip_loss_one_rank(r: int, image_features, text_features, logit_scale, logit_bias) -> torch.Tensor:
    loss = siglip_block_loss(image_features[r], text_features[r], logit_scale, logit_bias)
    for s in range(len(image_features)):
        if s != r:
            loss = loss + siglip_block_loss(image_features[r], text_features[s], logit_scale, logit_bias, True)
    return loss


def siglip_loss_ranks(image_features: Sequence[torch.Tensor], text_features: Sequence[torch.Tensor],
                      logit_scale, logit_bias) -> List[torch.Tensor]:
    """Per-rank SigLipLoss values (loss.py:406-489).  Every dist_impl visits each other rank's
    text block exactly once as a negative_only block, so the value is impl-independent
    (SURVEY §8c probe: all four impls agree to 8 digits)."""
    return [_siglip_loss_one_rank(r, image_features, text_features, logit_scale, logit_bias)
            for r in range(len(image_features))]


def siglip_loss_rank_grads(image_features, text_features, logit_scale, logit_bias, rank: int):
    """(loss_rank, d image_features[rank], d text_features[rank], d logit_scale, d logit_bias) as ONE rank of the
    reference sees them: feature gradients of the sum of all ranks' losses (the exchanges carry autograd,
    loss.py:226-311), scalar gradients of this rank's loss."""
    W = len(image_features)
    img = [f.detach() for f in image_features]
    txt = [f.detach() for f in text_features]
    img[rank] = img[rank].clone().requires_grad_(True)
    txt[rank] = txt[rank].clone().requires_grad_(True)
    out = None
    for r in range(W):
        scale = logit_scale.detach().clone().requires_grad_(True)
        bias = logit_bias.detach().clone().requires_grad_(True)
        loss = _siglip_loss_one_rank(r, img, txt, scale, bias)
        loss.backward()
        if r == rank:
            out = (loss.detach(), scale.grad.detach(), bias.grad.detach())
    return out[0], img[rank].grad, txt[rank].grad, out[1], out[2]


# --------------------------------------------------------------------------------------
# synthetic batch + a CPU train step (the `cpu_baseline` / `--impl reference` workload)
# --------------------------------------------------------------------------------------
def synthetic_batch(cfg: ClipCfg, batch: int, seed: int, dtype=torch.float32):
    """SURVEY §8d inputs: image ~ N(0,1), text uniform in [1, vocab-2], last token = vocab-1 (EOT = max id)."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, cfg.image_size, cfg.image_size, generator=g).to(dtype)
    text = torch.randint(1, cfg.t_vocab - 1, (batch, cfg.t_ctx), generator=g)
    text[:, -1] = cfg.t_vocab - 1
    return image, text


class CpuTrainer:
    """Reference train step restated: CLIPTask.training_forward (task/clip_task.py:41-46) +
    _make_train_step_no_accum_no_scaler (open_clip_train/train.py:163-185) + AdamW with the
    reference ViT defaults (open_clip_train/params.py:5-9,285) + clamp_logit_scale
    (task/image_text_task.py:91-101).  fp32, autograd through the functional forward above."""

    def __init__(self, cfg: ClipCfg, seed: int = 0, lr: float = 5e-4, wd: float = 0.2):
        self.cfg = cfg
        self.params = {k: v.requires_grad_(True) for k, v in init_params(cfg, seed).items()}
        # optim.py:67-75: no weight decay for ndim<=1 and no_weight_decay() names
        no_wd_names = {"positional_embedding", "visual.positional_embedding", "visual.class_embedding"}
        decay = [v for k, v in self.params.items() if v.ndim > 1 and k not in no_wd_names]
        no_decay = [v for k, v in self.params.items() if not (v.ndim > 1 and k not in no_wd_names)]
        self.opt = torch.optim.AdamW(
            [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": wd}],
            lr=lr, betas=(0.9, 0.98), eps=1e-6)

    def step(self, image: torch.Tensor, text: torch.Tensor) -> float:
        self.opt.zero_grad(set_to_none=True)
        out = clip_forward(self.params, self.cfg, image, text)
        loss = clip_loss(out["image_features"], out["text_features"], out["logit_scale"])
        loss.backward()
        self.opt.step()
        with torch.no_grad():
            self.params["logit_scale"].clamp_(0, math.log(100))
        return float(loss.detach())
