"""NativeCLIP — drop-in for the reference `CLIP` module (model.py:318-548) on B200.

Same constructor surface that matters to callers (embed_dim, vision_cfg, text_cfg, init_logit_scale,
init_logit_bias, output_dict), same parameter names / shapes / dtypes as `--precision bf16`
(convert_weights_to_lp model.py:738-765: Linear/conv/in_proj/projection weights bf16, LayerNorm affine,
embeddings and logit_scale fp32) so checkpoints round-trip, same forward/encode_* contracts
 — but every FLOP of the towers runs in libclipn.so (sm_100a kernels) and the backward is
scheduled by hand in tower.py (one autograd.Function per tower).

There is no CPU path: calling this on CPU tensors raises.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops, tower
from ._lib import ClipnError

BF16, F32 = torch.bfloat16, torch.float32


def _is_lowp(name: str) -> bool:
    """convert_weights_to_lp (model.py:738-765) targets."""
    return name in ("text_projection", "visual.proj", "visual.conv1.weight") or ".attn." in name or ".mlp." in name


class _Node(nn.Module):
    """Pure parameter container used to reproduce the reference's module tree / state_dict names."""


class _VisualNode(_Node):
    def __init__(self, image_size: int):
        super().__init__()
        self.image_size = (image_size, image_size)  # image_text_task.py:44 reads visual.image_size

    def no_weight_decay(self):
        return {"positional_embedding", "class_embedding"}  # transformer.py:778-781

    def layer_groups(self, pooler_in_head: bool = True):
        """Input -> output partition shared by lock() and layer-wise LR decay (transformer.py:718-743):
        embeddings | layer.i (last block together with ln_post) | proj."""
        groups = [("embeddings", [self.conv1, self.class_embedding, self.positional_embedding, self.ln_pre])]
        blocks = list(self.transformer.resblocks)
        for i, blk in enumerate(blocks):
            groups.append((f"layer.{i}", [blk, self.ln_post] if i == len(blocks) - 1 else [blk]))
        groups.append(("proj", [self.proj]))
        return groups

    def lock(self, unlocked_groups: int = 0, freeze_bn_stats: bool = False):
        _lock_groups(self.layer_groups(), unlocked_groups)  # transformer.py:745-753


def _lock_groups(groups, unlocked: int):
    """Freeze bottom-up, leave the top `unlocked` groups trainable; every group is set explicitly so repeated calls
    with different counts are idempotent (transformer.py:2044-2054)."""
    n_freeze = len(groups) if not unlocked else len(groups) - unlocked
    for i, (_, members) in enumerate(groups):
        for m in members:
            for p in ([m] if isinstance(m, nn.Parameter) else m.parameters()):
                p.requires_grad = i >= n_freeze


def _block_node(d: int, mlp: int) -> _Node:
    b = _Node()
    # registration order = the reference's (ln_1, attn, ln_2, mlp): named_parameters() order is what index-based
    # optimizer checkpoints and DDP buckets follow
    b.ln_1, b.attn, b.ln_2, b.mlp = _Node(), _Node(), _Node(), _Node()
    b.attn.out_proj, b.mlp.c_fc, b.mlp.c_proj = _Node(), _Node(), _Node()
    for ln in (b.ln_1, b.ln_2):
        ln.weight = nn.Parameter(torch.ones(d))
        ln.bias = nn.Parameter(torch.zeros(d))
    b.attn.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
    b.attn.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
    b.attn.out_proj.weight = nn.Parameter(torch.empty(d, d))
    b.attn.out_proj.bias = nn.Parameter(torch.zeros(d))
    b.mlp.c_fc.weight = nn.Parameter(torch.empty(mlp, d))
    b.mlp.c_fc.bias = nn.Parameter(torch.zeros(mlp))
    b.mlp.c_proj.weight = nn.Parameter(torch.empty(d, mlp))
    b.mlp.c_proj.bias = nn.Parameter(torch.zeros(d))
    return b


def _transformer_node(d: int, layers: int, mlp_ratio: float) -> _Node:
    t = _Node()
    t.resblocks = nn.ModuleList([_block_node(d, int(d * mlp_ratio)) for _ in range(layers)])
    return t


class _TowerFn(torch.autograd.Function):
    """One tower = one autograd node. forward(inputs, *params) -> features; backward -> every param grad."""

    @staticmethod
    def forward(ctx, model: "NativeCLIP", which: str, normalize: bool, need_grad: bool, inp: torch.Tensor, *params):
        names = model._tower_param_names[which]
        P = dict(zip(names, params))
        cfg = model._vcfg if which == "visual" else model._tcfg
        fwd = tower.vision_forward if which == "visual" else tower.text_forward
        feat, saved = fwd(P, cfg, inp, normalize, model._scratch[which], need_grad,
                          checkpoint=bool(model.grad_checkpointing))
        ctx.model, ctx.which, ctx.saved, ctx.P, ctx.names = model, which, saved, P, names
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        model, which, P, names = ctx.model, ctx.which, ctx.P, ctx.names
        if ctx.saved is None:
            raise ClipnError("backward through a NativeCLIP tower that ran without saved activations")
        cfg = model._vcfg if which == "visual" else model._tcfg
        arena = model._grad_arena(which)
        arena["flat32"].zero_()
        G = arena["views32"]
        # The fp32 arena is this node's ACCUMULATOR only; what autograd receives are views of two buffers owned by
        # this call: one bf16 cast of the low-precision segment, one copy of the fp32 segment.  The arena can therefore
        # be zeroed and refilled by the next backward of this tower — another micro-batch (`--accum-freq`,
        # train.py:236-311), or a second use of the tower in the SAME graph (multi-view / multi-caption losses) whose
        # first gradients still sit in the engine's input buffers.
        n_lowp = arena["n_lowp"]
        out16 = torch.empty_like(arena["flat16"]) if n_lowp else None
        out32 = torch.empty_like(arena["flat32"][n_lowp:])
        sync = model._grad_sync

        def finalize(lo16, hi16, lo32, hi32):
            """arena slices -> the per-call gradient buffers (+ asynchronous averaging all-reduce under grad sync)"""
            if hi16 > lo16:
                ops.cast_f32_to_bf16(arena["flat32"][lo16:hi16], out=out16[lo16:hi16])
                if sync is not None:
                    sync.reduce(out16[lo16:hi16])
            if hi32 > lo32:
                out32[lo32 - n_lowp:hi32 - n_lowp].copy_(arena["flat32"][lo32:hi32])
                if sync is not None:
                    sync.reduce(out32[lo32 - n_lowp:hi32 - n_lowp])

        done16, done32 = [], []
        if sync is not None:
            spans = arena["block_spans"]

            def on_block(pre):
                lo16, hi16, lo32, hi32 = spans[pre]
                finalize(lo16, hi16, lo32, hi32)
                done16.append((lo16, hi16))
                done32.append((lo32, hi32))
            ctx.saved.extra["on_block_grads_ready"] = on_block
        bwd = tower.vision_backward if which == "visual" else tower.text_backward
        bwd(P, G, cfg, ctx.saved, dfeat, model._scratch[which])
        ctx.saved = None
        # whatever the block callbacks did not cover (embeddings, head; everything without grad sync): the gaps
        for (seg_lo, seg_hi, done, is16) in ((0, n_lowp, done16, True), (n_lowp, arena["flat32"].numel(), done32, False)):
            cur = seg_lo
            for lo, hi in sorted(done) + [(seg_hi, seg_hi)]:
                if lo > cur:
                    finalize(cur, lo, 0, 0) if is16 else finalize(0, 0, cur, lo)
                cur = max(cur, hi)
        if sync is not None:
            sync.wait()
        grads = []
        for n, p in zip(names, ctx.P.values()):
            if not p.requires_grad:
                grads.append(None)
                continue
            off, k = arena["offsets"][n]
            src = out16[off:off + k] if p.dtype == BF16 else out32[off - n_lowp:off - n_lowp + k]
            grads.append(src.view(p.shape))
        return (None, None, None, None, None, *grads)


class GradSync:
    """Data-parallel gradient averaging driven by the towers' own backward (replaces DistributedDataParallel,
    base_task.py:227, for NativeCLIP): every residual block's slice of the flat gradient buffers is all-reduced
    (NCCL, asynchronously, on NCCL's stream) the moment that block's backward has finished, so the exchange overlaps the
    remaining backward instead of starting when a whole tower's autograd node returns.  Scalars outside the towers
    (logit_scale, logit_bias) are averaged by post-accumulate hooks.  Average = DDP's convention (sum / world)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.avg = dist.get_backend(self.group) == "nccl"  # gloo has no AVG: sum, then scale
        self.pending = []

    def reduce(self, t: torch.Tensor):
        if self.world == 1:
            return
        op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
        self.pending.append((self.dist.all_reduce(t, op=op, group=self.group, async_op=True), t))

    def wait(self):
        for work, t in self.pending:
            work.wait()
            if not self.avg:
                t.div_(self.world)
        self.pending.clear()


class NativeCLIP(nn.Module):
    def __init__(self, embed_dim: int, vision_cfg: dict, text_cfg: dict, quick_gelu: bool = False,
                 init_logit_scale: float = math.log(1 / 0.07), init_logit_bias: Optional[float] = None,
                 output_dict: bool = False, device="cuda"):
        super().__init__()
        if quick_gelu:
            raise ClipnError("NativeCLIP implements nn.GELU (erf); quick_gelu configs are out of scope")
        v, t = dict(vision_cfg), dict(text_cfg)
        head_width = v.get("head_width", 64)
        if head_width != 64 or t["width"] // t["heads"] != 64:
            raise ClipnError("NativeCLIP kernels are specialised for head_dim 64 (every native open_clip ViT)")
        self.output_dict = output_dict
        self.embed_dim = embed_dim
        self.context_length = t.get("context_length", 77)
        self.vocab_size = t.get("vocab_size", 49408)
        vw, tw = v["width"], t["width"]
        if isinstance(v["image_size"], (tuple, list)):  # CLIPVisionCfg.image_size may be an (h, w) pair (model.py:41)
            if len(set(v["image_size"])) != 1:
                raise ClipnError(f"NativeCLIP supports square images only, got image_size={v['image_size']}")
            v["image_size"] = int(v["image_size"][0])
        if isinstance(v["patch_size"], (tuple, list)):
            if len(set(v["patch_size"])) != 1:
                raise ClipnError(f"NativeCLIP supports square patches only, got patch_size={v['patch_size']}")
            v["patch_size"] = int(v["patch_size"][0])
        grid = v["image_size"] // v["patch_size"]
        mlp_ratio = v.get("mlp_ratio", 4.0)
        self._vcfg = tower.TowerCfg(width=vw, layers=v["layers"], heads=vw // 64, seq=grid * grid + 1, causal=False,
                                    prefix="visual.transformer", embed_dim=embed_dim, image_size=v["image_size"],
                                    patch=v["patch_size"])
        self._tcfg = tower.TowerCfg(width=tw, layers=t["layers"], heads=t["heads"], seq=self.context_length, causal=True,
                                    prefix="transformer", embed_dim=embed_dim, vocab=self.vocab_size)

        # ---- parameter tree with the reference's names
        self.visual = _VisualNode(v["image_size"])
        self.visual.conv1 = _Node()
        self.visual.conv1.weight = nn.Parameter(torch.empty(vw, 3, v["patch_size"], v["patch_size"]))
        self.visual.class_embedding = nn.Parameter(torch.empty(vw))
        self.visual.positional_embedding = nn.Parameter(torch.empty(grid * grid + 1, vw))
        # submodule order as in the reference: conv1, ln_pre, transformer, ln_post (named_parameters() order)
        self.visual.ln_pre = _Node()
        self.visual.transformer = _transformer_node(vw, v["layers"], mlp_ratio)
        self.visual.ln_post = _Node()
        for ln in (self.visual.ln_pre, self.visual.ln_post):
            ln.weight = nn.Parameter(torch.ones(vw))
            ln.bias = nn.Parameter(torch.zeros(vw))
        self.visual.proj = nn.Parameter(torch.empty(vw, embed_dim))
        self.transformer = _transformer_node(tw, t["layers"], t.get("mlp_ratio", 4.0))
        self.token_embedding = _Node()
        self.token_embedding.weight = nn.Parameter(torch.empty(self.vocab_size, tw))
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, tw))
        self.ln_final = _Node()
        self.ln_final.weight = nn.Parameter(torch.ones(tw))
        self.ln_final.bias = nn.Parameter(torch.zeros(tw))
        self.text_projection = nn.Parameter(torch.empty(tw, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * init_logit_scale)
        self.logit_bias = nn.Parameter(torch.ones([]) * init_logit_bias) if init_logit_bias is not None else None
        # causal mask is a predicate inside the attention kernel; keep the (non-persistent) buffer for API parity
        self.register_buffer("attn_mask", torch.full((self.context_length, self.context_length), float("-inf")).triu_(1),
                             persistent=False)

        self.init_parameters()
        self._apply_dtype_contract()
        self.to(device)
        self._finalize()

    # ------------------------------------------------------------------ construction helpers
    @torch.no_grad()
    def init_parameters(self):
        """Reference init distributions (transformer.py:145-155,1664-1685,641-645,719; model.py:326)."""
        vw, tw = self._vcfg.width, self._tcfg.width
        vs = vw ** -0.5
        nn.init.normal_(self.visual.class_embedding, std=vs)
        nn.init.normal_(self.visual.positional_embedding, std=vs)
        nn.init.normal_(self.visual.proj, std=vs)
        nn.init.kaiming_uniform_(self.visual.conv1.weight, a=math.sqrt(5))
        for blk in self.visual.transformer.resblocks:  # vision tower keeps torch defaults (+ MHA-style attn init)
            nn.init.xavier_uniform_(blk.attn.in_proj_weight)
            for lin in (blk.attn.out_proj, blk.mlp.c_fc, blk.mlp.c_proj):
                nn.init.kaiming_uniform_(lin.weight, a=math.sqrt(5))
            for lin in (blk.mlp.c_fc, blk.mlp.c_proj):
                bound = 1 / math.sqrt(lin.weight.shape[1])
                nn.init.uniform_(lin.bias, -bound, bound)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        L = self._tcfg.layers
        proj_std, attn_std, fc_std = (tw ** -0.5) * ((2 * L) ** -0.5), tw ** -0.5, (2 * tw) ** -0.5
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)
            for lin in (blk.mlp.c_fc, blk.mlp.c_proj):
                bound = 1 / math.sqrt(lin.weight.shape[1])
                nn.init.uniform_(lin.bias, -bound, bound)
        nn.init.normal_(self.text_projection, std=tw ** -0.5)

    def _apply_dtype_contract(self):
        for name, p in self.named_parameters():
            p.data = p.data.to(BF16 if _is_lowp(name) else F32)

    def _finalize(self):
        names = [n for n, _ in self.named_parameters()]
        self._tower_param_names = {
            "visual": [n for n in names if n.startswith("visual.")],
            "text": [n for n in names if not n.startswith("visual.") and n not in ("logit_scale", "logit_bias")],
        }
        self._scratch = {"visual": tower.Scratch(), "text": tower.Scratch()}
        self._arenas: Dict[str, dict] = {}
        self.grad_checkpointing = False
        self._grad_sync: Optional[GradSync] = None

    def _grad_arena(self, which: str) -> dict:
        """Flat fp32 gradient accumulators for one tower (low-precision params first)."""
        a = self._arenas.get(which)
        params = dict(self.named_parameters())
        names = self._tower_param_names[which]
        dev = params[names[0]].device
        if a is None or a["flat32"].device != dev:
            lowp = [n for n in names if params[n].dtype == BF16]
            highp = [n for n in names if params[n].dtype != BF16]
            pad = lambda k: (k + 63) // 64 * 64  # keep every view 256-byte aligned (TMA needs 16)
            n_lowp = sum(pad(params[n].numel()) for n in lowp)
            total = n_lowp + sum(pad(params[n].numel()) for n in highp)
            flat32 = torch.zeros(total, dtype=F32, device=dev)
            flat16 = torch.empty(n_lowp, dtype=BF16, device=dev)
            v32, offsets, off = {}, {}, 0
            for n in lowp + highp:
                k = params[n].numel()
                v32[n] = flat32[off:off + k].view(params[n].shape)
                offsets[n] = (off, k)
                off += pad(k)
            # per residual block: [lo, hi) of its parameters inside the bf16 segment and inside the fp32 segment
            # (named_parameters() order keeps a block's tensors together in each segment)
            spans = {}
            for n in lowp + highp:
                if ".resblocks." not in n:
                    continue
                pre = n[:n.index(".resblocks.") + len(".resblocks.")] + n.split(".resblocks.")[1].split(".")[0]
                o, k = offsets[n]
                sp = spans.setdefault(pre, [n_lowp, 0, total, n_lowp])
                if n in lowp:
                    sp[0], sp[1] = min(sp[0], o), max(sp[1], o + pad(k))
                else:
                    sp[2], sp[3] = min(sp[2], o), max(sp[3], o + pad(k))
            for pre, sp in spans.items():
                sp[1], sp[3] = min(sp[1], n_lowp), min(sp[3], total)
            # flat16 is only the shape/dtype template of the per-call bf16 gradient buffer (never written)
            a = {"flat32": flat32, "flat16": flat16, "views32": v32, "offsets": offsets, "n_lowp": n_lowp,
                 "block_spans": {k: tuple(v) for k, v in spans.items()}}
            self._arenas[which] = a
        return a

    def enable_grad_sync(self, group=None) -> "GradSync":
        """Average gradients across the data-parallel group from inside the backward (per-block overlap) INSTEAD of
        wrapping the module in DistributedDataParallel.  Parameters must already be identical on all ranks."""
        self._grad_sync = GradSync(group)
        sync = self._grad_sync
        for p in (self.logit_scale, self.logit_bias):
            if p is not None:
                def hook(param, sync=sync):
                    sync.reduce(param.grad)
                    sync.wait()
                p.register_post_accumulate_grad_hook(hook)
        return sync

    # ------------------------------------------------------------------ reference API surface
    def set_grad_checkpointing(self, enable: bool = True, impl: str = "inline"):
        # model.py:377-379 / transformer.py:397-402: keep only each block's input and re-run the block in the
        # backward (d bf16 per token per block instead of 10-16*d; one extra block forward per block).
        self.grad_checkpointing = enable

    def lock_image_tower(self, unlocked_groups: int = 0, freeze_bn_stats: bool = False):
        # model.py:368-370 -> VisionTransformer.lock: unlocked_groups counts the top groups (proj first) left trainable
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats)

    def text_layer_groups(self, pooler_in_head: bool = True):
        """embeddings | layer.i (last block together with ln_final) | proj (transformer.py:1999-2031)."""
        groups = [("embeddings", [self.token_embedding, self.positional_embedding])]
        blocks = list(self.transformer.resblocks)
        for i, blk in enumerate(blocks):
            groups.append((f"layer.{i}", [blk, self.ln_final] if i == len(blocks) - 1 else [blk]))
        groups.append(("proj", [self.text_projection]))
        return groups

    def lock_text_tower(self, unlocked_layers: int = 0, freeze_layer_norm: bool = True, pooler_in_head: bool = True):
        _lock_groups(self.text_layer_groups(pooler_in_head), unlocked_layers)  # transformer.py:2056-2076

    def no_weight_decay(self):
        # model.py:381-387
        return {"positional_embedding", "visual.positional_embedding", "visual.class_embedding"}

    def _run_tower(self, which: str, inp: torch.Tensor, normalize: bool) -> torch.Tensor:
        if not inp.is_cuda:
            raise ClipnError("NativeCLIP runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        params = dict(self.named_parameters())
        plist = [params[n] for n in self._tower_param_names[which]]
        # grad mode is always off inside Function.forward, so decide here whether activations must be saved
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in plist)
        return _TowerFn.apply(self, which, normalize, need_grad, inp, *plist)

    def encode_image(self, image, normalize: bool = False):
        return self._run_tower("visual", image, normalize)

    def encode_text(self, text, normalize: bool = False):
        return self._run_tower("text", text, normalize)

    def forward(self, image: Optional[torch.Tensor] = None, text: Optional[torch.Tensor] = None):
        # model.py:528-548
        image_features = self.encode_image(image, normalize=True) if image is not None else None
        text_features = self.encode_text(text, normalize=True) if text is not None else None
        if self.output_dict:
            out = {"image_features": image_features, "text_features": text_features,
                   "logit_scale": self.logit_scale.exp()}
            if self.logit_bias is not None:
                out["logit_bias"] = self.logit_bias.clone()
            return out
        if self.logit_bias is not None:
            return image_features, text_features, self.logit_scale.exp(), self.logit_bias.clone()
        return image_features, text_features, self.logit_scale.exp()

    # ------------------------------------------------------------------ checkpoints
    @torch.no_grad()
    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Load a reference CLIP state_dict (fp32 or bf16); values are cast to this module's dtype contract."""
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own and k != "attn_mask"]
        if missing or unexpected:
            raise ClipnError(f"state_dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
        for k, p in own.items():
            p.copy_(sd[k].to(device=p.device, dtype=p.dtype).view(p.shape))


CONFIGS = {
    "ViT-B-32": dict(embed_dim=512, vision_cfg=dict(image_size=224, layers=12, width=768, patch_size=32),
                     text_cfg=dict(context_length=77, vocab_size=49408, width=512, heads=8, layers=12)),
    "ViT-B-16": dict(embed_dim=512, vision_cfg=dict(image_size=224, layers=12, width=768, patch_size=16),
                     text_cfg=dict(context_length=77, vocab_size=49408, width=512, heads=8, layers=12)),
    # model_configs/ViT-L-14-336.json (577 tokens: long-sequence attention kernels; patch 14: K-padded im2row)
    "ViT-L-14-336": dict(embed_dim=768, vision_cfg=dict(image_size=336, layers=24, width=1024, patch_size=14),
                         text_cfg=dict(context_length=77, vocab_size=49408, width=768, heads=12, layers=12)),
    "tiny": dict(embed_dim=128, vision_cfg=dict(image_size=64, layers=2, width=128, patch_size=16),
                 text_cfg=dict(context_length=20, vocab_size=512, width=128, heads=2, layers=2)),
}


def create_model(name: str, output_dict: bool = True, device="cuda", **kw) -> NativeCLIP:
    """Counterpart of open_clip.create_model(name, precision='bf16', output_dict=True) (factory.py:264)."""
    cfg = CONFIGS[name]
    return NativeCLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"], output_dict=output_dict, device=device, **kw)
