"""Tensor-level wrappers over the C ABI (one function per libclipn entry point).

torch is plumbing here: it owns device memory and the stream; every FLOP runs in libclipn.so.
All functions enqueue on torch.cuda.current_stream() and never synchronise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L

BF16 = torch.bfloat16
F32 = torch.float32


# instrumentation used by bench.py: kernel-launch counter and CUDA events around one GEMM signature
LAUNCHES = 0
PROFILE_KEY = None          # (M, N, K, epilogue), "all", or None
PROFILE_EVENTS = []         # [(start_event, end_event, (M, N, K, epilogue, a_mn, b_mn))] on the launching stream


def _call(rc: int, n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n
    L.check(rc)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise L.ClipnError(f"{name}: expected a CUDA tensor (open_clip_b200 has no CPU path)")
    if t.dtype != dtype:
        raise L.ClipnError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise L.ClipnError(f"{name}: expected a contiguous tensor")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, epilogue: int = L.EPI_STORE,
         bias: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         out2: Optional[torch.Tensor] = None, alpha: float = 1.0, splits: int = 1, ref: bool = False,
         **extra) -> torch.Tensor:
    """C[M,N] = epilogue(alpha * A . B^T); see include/clipn.h for operand layouts.
    a: [M,K] (a_mn False) or [K,M] (a_mn True); b: [N,K] (b_mn False) or [K,N] (b_mn True)."""
    _chk(a, BF16, "gemm.a")
    _chk(b, BF16, "gemm.b")
    assert a.dim() == 2 and b.dim() == 2
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise L.ClipnError(f"gemm: K mismatch {K} vs {Kb}")
    f32_out = epilogue in (L.EPI_ACCUM_F32, L.EPI_STORE_F32)
    if out is None and epilogue not in (L.EPI_LSE,):
        out = torch.empty((M, N), dtype=F32 if f32_out else BF16, device=a.device)
    d = L.GemmDesc()
    d.a, d.lda, d.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn)
    d.b, d.ldb, d.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn)
    if out is not None:
        _chk(out, F32 if f32_out else BF16, "gemm.out")
        assert tuple(out.shape) == (M, N)
        d.c, d.ldc = out.data_ptr(), out.stride(0)
    if out2 is not None:
        _chk(out2, BF16, "gemm.out2")
        d.c2, d.ldc2 = out2.data_ptr(), out2.stride(0)
    if bias is not None:
        _chk(bias, BF16, "gemm.bias")
        d.bias = bias.data_ptr()
    if aux is not None:
        _chk(aux, BF16, "gemm.aux")
        d.aux, d.ldaux = aux.data_ptr(), aux.stride(0)
    d.m, d.n, d.k = M, N, K
    d.epilogue, d.alpha, d.splits = epilogue, alpha, splits
    for k, v in extra.items():
        setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    fn = L.lib().clipn_gemm_ref if ref else L.lib().clipn_gemm
    if PROFILE_KEY is not None and (PROFILE_KEY == "all" or PROFILE_KEY == (M, N, K, epilogue)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _call(fn(C.byref(d), _stream()))
        e1.record()
        PROFILE_EVENTS.append((e0, e1, (M, N, K, epilogue, bool(a_mn), bool(b_mn))))
    else:
        _call(fn(C.byref(d), _stream()))
    return out


_SMS: Optional[int] = None


def _sm_count() -> int:
    """SMs of the current device (clipn_device_info); 148 (B200) where no GPU is visible — the CPU test container."""
    global _SMS
    if _SMS is None:
        _SMS = 148
        if torch.cuda.is_available():
            sms, major, minor = C.c_int(0), C.c_int(0), C.c_int(0)
            if L.lib().clipn_device_info(C.byref(sms), C.byref(major), C.byref(minor)) == 0 and sms.value > 0:
                _SMS = sms.value
    return _SMS


def wgrad_splits(m_out: int, n_out: int, k: int) -> int:
    """split-K factor for a weight-gradient GEMM: minimises rounds x (k-blocks per item + fixed cost per item), the
    fixed cost (accumulator drain + fp32 TMA reduce-add of the tile, pipeline fill) taken as 32 k-blocks.
    (Maximising SM utilisation alone picked 30 splits for the [2304, 768] in_proj gradient — 0.995 utilisation, but
    11 short rounds and 30 reduce-adds per tile: measured 0.64 of peak against 0.92-0.95 for its 2-split siblings.)"""
    bn = L.lib().clipn_gemm_tile_n(n_out)
    tiles = ((m_out + 127) // 128) * ((n_out + bn - 1) // bn)
    kblocks = (k + 63) // 64
    sms = _sm_count()
    best, best_cost = 1, float("inf")
    for s in range(1, 33):
        if s > 1 and kblocks // s < 16:   # keep >= 16 k-blocks (1024 rows) per work item
            break
        rounds = (tiles * s + sms - 1) // sms
        cost = rounds * ((kblocks + s - 1) // s + 32)
        if cost < best_cost:
            best, best_cost = s, cost
    return best


def layernorm_fwd(x, gamma, beta, out=None, eps: float = 1e-5, save_stats: bool = True):
    _chk(x, BF16, "ln.x"); _chk(gamma, F32, "ln.gamma"); _chk(beta, F32, "ln.beta")
    rows, d = x.shape
    y = out if out is not None else torch.empty_like(x)
    mean = torch.empty(rows, dtype=F32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=F32, device=x.device) if save_stats else None
    _call(L.lib().clipn_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(mean),
                                        _ptr(rstd), rows, d, eps, _stream()))
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, resid=None, out=None, resid_sum=None):
    """resid_sum (fp32 [d], +=): column sums of `resid` = bias gradient of the Linear that fed the residual stream."""
    _chk(dy, BF16, "lnb.dy"); _chk(x, BF16, "lnb.x")
    rows, d = x.shape
    dx = out if out is not None else torch.empty_like(x)
    if resid_sum is not None:
        _chk(resid_sum, F32, "lnb.resid_sum")
    _call(L.lib().clipn_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                        _ptr(resid), dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), _ptr(resid_sum), rows, d,
                                        _stream()))
    return dx


def attention_fwd(qkv, batch, seq, heads, causal, out=None):
    _chk(qkv, BF16, "attn.qkv")
    d = heads * 64
    assert tuple(qkv.shape) == (batch * seq, 3 * d)
    o = out if out is not None else torch.empty((batch * seq, d), dtype=BF16, device=qkv.device)
    lse = torch.empty((batch, heads, seq), dtype=F32, device=qkv.device)
    _call(L.lib().clipn_attention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), batch, seq, heads, int(causal),
                                        64 ** -0.5, _stream()))
    return o, lse


def attention_bwd(qkv, o, do, lse, batch, seq, heads, causal, out=None, dbias=None):
    """dbias (fp32 [3d], +=): in_proj_bias gradient = column sums of dqkv, fused into the kernel."""
    _chk(qkv, BF16, "attnb.qkv"); _chk(o, BF16, "attnb.o"); _chk(do, BF16, "attnb.do"); _chk(lse, F32, "attnb.lse")
    dqkv = out if out is not None else torch.empty_like(qkv)
    if dbias is not None:
        _chk(dbias, F32, "attnb.dbias")
    _call(L.lib().clipn_attention_bwd(qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                        _ptr(dbias), batch, seq, heads, int(causal), 64 ** -0.5, _stream()))
    return dqkv


def patch_cols(chans, patch):
    """Row length of the im2row matrix: chans*patch*patch rounded up to 8 (TMA wants 16-byte row pitches)."""
    return (chans * patch * patch + 7) // 8 * 8


def patchify(image, patch):
    """im2row for conv1. Returns [B*gh*gw, patch_cols]; columns beyond chans*patch*patch (patch 14 only) are zero."""
    _chk(image, BF16, "patchify.image")
    B, Cc, H, W = image.shape
    k = Cc * patch * patch
    if patch % 8 == 0:
        out = torch.empty((B * (H // patch) * (W // patch), k), dtype=BF16, device=image.device)
        _call(L.lib().clipn_patchify(image.data_ptr(), out.data_ptr(), B, Cc, H, W, patch, _stream()))
        return out
    ld = patch_cols(Cc, patch)
    out = torch.empty((B * (H // patch) * (W // patch), ld), dtype=BF16, device=image.device)
    _call(L.lib().clipn_patchify_padded(image.data_ptr(), out.data_ptr(), ld, B, Cc, H, W, patch, _stream()))
    return out


def accum_rows_f32(dst, src, cols):
    """dst[:, :cols] += src[:, :cols] (fp32, row pitches from the tensors)."""
    _chk(dst, F32, "accum.dst"); _chk(src, F32, "accum.src")
    _call(L.lib().clipn_accum_rows_f32(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.stride(0), dst.shape[0], cols,
                                       _stream()))
    return dst


def vision_embed_fwd(patch_out, cls, pos, batch, npatch):
    _chk(patch_out, BF16, "vemb.patch_out"); _chk(cls, F32, "vemb.cls"); _chk(pos, F32, "vemb.pos")
    d = patch_out.shape[1]
    x = torch.empty((batch * (npatch + 1), d), dtype=BF16, device=patch_out.device)
    _call(L.lib().clipn_vision_embed_fwd(patch_out.data_ptr(), cls.data_ptr(), pos.data_ptr(), x.data_ptr(), batch,
                                           npatch, d, _stream()))
    return x


def vision_embed_bwd(dx, dcls, dpos, batch, npatch):
    _chk(dx, BF16, "vembb.dx")
    d = dx.shape[1]
    dpatch = torch.empty((batch * npatch, d), dtype=BF16, device=dx.device)
    _call(L.lib().clipn_vision_embed_bwd(dx.data_ptr(), dpatch.data_ptr(), _ptr(dcls), _ptr(dpos), batch, npatch, d,
                                           _stream()))
    return dpatch


def text_embed_fwd(ids, table, pos):
    _chk(ids, torch.int64, "temb.ids"); _chk(table, F32, "temb.table"); _chk(pos, F32, "temb.pos")
    B, S = ids.shape
    vocab, d = table.shape
    # nn.Embedding raises on an out-of-range id (tokenizer / vocab mismatch); the kernel would clamp it silently, so the
    # range is asserted on the device (asynchronously: no host sync, a violation surfaces as a CUDA device assert)
    torch._assert_async(((ids >= 0) & (ids < vocab)).all(), "open_clip_b200: token id outside [0, vocab_size)")
    x = torch.empty((B * S, d), dtype=BF16, device=ids.device)
    eot = torch.empty(B, dtype=torch.int32, device=ids.device)
    _call(L.lib().clipn_text_embed_fwd(ids.data_ptr(), table.data_ptr(), pos.data_ptr(), x.data_ptr(), eot.data_ptr(),
                                         B, S, d, vocab, _stream()))
    return x, eot


def text_embed_bwd(ids, dx, dtable, dpos):
    _chk(ids, torch.int64, "tembb.ids"); _chk(dx, BF16, "tembb.dx"); _chk(dtable, F32, "tembb.dtable")
    B, S = ids.shape
    vocab, d = dtable.shape
    _call(L.lib().clipn_text_embed_bwd(ids.data_ptr(), dx.data_ptr(), dtable.data_ptr(), _ptr(dpos), B, S, d, vocab,
                                         _stream()))


def gather_rows(x, idx, batch, seq):
    _chk(x, BF16, "gather.x")
    d = x.shape[1]
    out = torch.empty((batch, d), dtype=BF16, device=x.device)
    _call(L.lib().clipn_gather_rows(x.data_ptr(), _ptr(idx), out.data_ptr(), batch, seq, d, _stream()))
    return out


def scatter_rows(dpooled, idx, batch, seq, out=None):
    _chk(dpooled, BF16, "scatter.dpooled")
    d = dpooled.shape[1]
    dx = out if out is not None else torch.empty((batch * seq, d), dtype=BF16, device=dpooled.device)
    _call(L.lib().clipn_scatter_rows(dpooled.data_ptr(), _ptr(idx), dx.data_ptr(), batch, seq, d, _stream()))
    return dx


def l2norm_fwd(x):
    _chk(x, BF16, "l2.x")
    rows, d = x.shape
    y = torch.empty_like(x)
    inv = torch.empty(rows, dtype=F32, device=x.device)
    _call(L.lib().clipn_l2norm_fwd(x.data_ptr(), y.data_ptr(), inv.data_ptr(), rows, d, _stream()))
    return y, inv


def l2norm_bwd(dy, y, inv):
    assert dy.dtype in (F32, BF16) and dy.is_contiguous() and dy.is_cuda
    rows, d = y.shape
    dx = torch.empty_like(y)
    _call(L.lib().clipn_l2norm_bwd(dy.data_ptr(), int(dy.dtype == F32), y.data_ptr(), inv.data_ptr(), dx.data_ptr(),
                                     rows, d, _stream()))
    return dx


def colsum(x, out):
    _chk(x, BF16, "colsum.x"); _chk(out, F32, "colsum.out")
    rows, n = x.shape
    _call(L.lib().clipn_colsum(x.data_ptr(), x.stride(0), out.data_ptr(), rows, n, _stream()))
    return out


def cast_f32_to_bf16(x, out=None):
    _chk(x, F32, "cast.x")
    y = out if out is not None else torch.empty(x.shape, dtype=BF16, device=x.device)
    _call(L.lib().clipn_cast_f32_to_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()))
    return y


# ---------------------------------------------------------------------------- contrastive loss pieces
def _ptr_array(ptrs: Sequence[int]):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr


def peer_gemm_tile_n(world: int, b: int, e: int) -> int:
    """Column-tile width of the fused peer-streaming forward for this shape, 0 if it does not take it."""
    return int(L.lib().clipn_peer_gemm_tile_n(world, b, e))


def peer_gather(txt_ptrs: Sequence[int], img_ptrs: Sequence[int], b: int, e: int, gather_txt: torch.Tensor,
                gather_img: torch.Tensor):
    """The P2P gather step alone (every rank's [B,E] block -> local [W*B,E] copies)."""
    _call(L.lib().clipn_peer_gather(_ptr_array(txt_ptrs), _ptr_array(img_ptrs), len(txt_ptrs), b, e,
                                     gather_txt.data_ptr(), gather_img.data_ptr(), _stream()))


def _profiled(sig):
    """bench.py instrumentation: CUDA events on the launching stream around one C-ABI call (context manager)."""
    class _Ctx:
        def __enter__(self):
            self.on = PROFILE_KEY == "all"
            if self.on:
                self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *exc):
            if self.on and exc[0] is None:
                self.e1.record()
                PROFILE_EVENTS.append((self.e0, self.e1, sig))
    return _Ctx()


def stage_timing(enable: bool) -> None:
    """Measurement aid: record CUDA events between the gather / GEMM / combine stages of every fused ClipLoss forward."""
    _call(L.lib().clipn_stage_timing(1 if enable else 0))


def stage_times():
    """(calls, gather_ms, gemm_ms, combine_ms): mean stage times since stage_timing(True)."""
    buf = (C.c_float * 3)()
    n = L.lib().clipn_stage_times(buf)
    return int(n), float(buf[0]), float(buf[1]), float(buf[2])


def clip_fwd_fused(img: torch.Tensor, txt: torch.Tensor, txt_ptrs: Sequence[int], img_ptrs: Sequence[int], rank: int,
                   scale: torch.Tensor, gather_txt: Optional[torch.Tensor], gather_img: Optional[torch.Tensor]):
    """Both directions of the ClipLoss forward in one launch (+ a tiny combine kernel): returns lse [2, B] (image rows,
    text rows) and the 1-element loss accumulator (already the local cross-entropy value, loss.py:135-139).
    `*_ptrs`: one device pointer per rank to that rank's bf16 [B, E] features (peer-mapped).  `gather_*`: optional
    local bf16 [W*B, E] buffers that receive the gathered column operands as a by-product."""
    _chk(img, BF16, "fused.img"); _chk(txt, BF16, "fused.txt"); _chk(scale, F32, "fused.scale")
    b, e = img.shape
    world = len(txt_ptrs)
    ws = torch.empty(L.lib().clipn_clip_fwd_fused_workspace(world, b, e), dtype=F32, device=img.device)
    lse = torch.empty((2, b), dtype=F32, device=img.device)
    loss = torch.zeros(1, dtype=F32, device=img.device)
    # bench.py: the fused gather + logits + LSE GEMM of BOTH directions (and its combine kernel) as one timed signature
    with _profiled((2 * b, world * b, e, L.EPI_LSE, False, False)):
        _call(L.lib().clipn_clip_fwd_fused(img.data_ptr(), txt.data_ptr(), _ptr_array(txt_ptrs), _ptr_array(img_ptrs),
                                             world, rank, b, e, 1.0, scale.data_ptr(), _ptr(gather_txt),
                                             _ptr(gather_img), lse.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                             _stream()), 2 + int(world > 1))
    return lse, loss


def siglip_fwd_fused(img, txt, txt_ptrs, img_ptrs, rank, scale, bias, gscale, gather_txt, gather_img, loss_acc,
                     scalar_acc, want_grad: bool):
    """SigLipLoss forward in one launch; with want_grad also d(logits) of both directions (bf16 [B, ld])."""
    _chk(img, BF16, "siglip.img"); _chk(txt, BF16, "siglip.txt")
    b, e = img.shape
    world = len(txt_ptrs)
    n = world * b
    ld = (n + 7) // 8 * 8
    dl_i = torch.empty((b, ld), dtype=BF16, device=img.device) if want_grad else None
    dl_t = torch.empty((b, ld), dtype=BF16, device=img.device) if want_grad else None
    with _profiled(((2 if want_grad else 1) * b, n, e, L.EPI_SIGLIP, False, False)):
        _call(L.lib().clipn_siglip_fwd_fused(img.data_ptr(), txt.data_ptr(), _ptr_array(txt_ptrs), _ptr_array(img_ptrs),
                                               world, rank, b, e, scale.data_ptr(), bias.data_ptr(), gscale,
                                               _ptr(gather_txt), _ptr(gather_img), loss_acc.data_ptr(),
                                               _ptr(scalar_acc), _ptr(dl_i), _ptr(dl_t), ld, _stream()),
              1 + int(world > 1))
    return dl_i, dl_t


def clip_lse_fwd(rows: torch.Tensor, cols: torch.Tensor, scale: torch.Tensor, label_offset: int):
    """Generic (any E % 8 == 0, n % 8 == 0) row log-sum-exp and label logit of scale * rows @ cols.T; cols is one
    LOCAL [n, E] tensor.  `scale` is a 1-element fp32 DEVICE tensor (no host sync)."""
    _chk(rows, BF16, "lse.rows"); _chk(cols, BF16, "lse.cols"); _chk(scale, F32, "lse.scale")
    m, e = rows.shape
    n = cols.shape[0]
    ws = torch.empty(L.lib().clipn_clip_lse_workspace(m, n), dtype=F32, device=rows.device)
    lse = torch.empty(m, dtype=F32, device=rows.device)
    pos = torch.zeros(m, dtype=F32, device=rows.device)
    with _profiled((m, n, e, L.EPI_LSE, False, False)):
        _call(L.lib().clipn_clip_lse_fwd(rows.data_ptr(), cols.data_ptr(), m, n, e, 1.0, scale.data_ptr(), label_offset,
                                           lse.data_ptr(), pos.data_ptr(), ws.data_ptr(), _stream()), 2)
    return lse, pos


def clip_dlogits(rows, cols, scale, label_offset, row_lse, col_lse, col_w, gscale, scalar_acc, row_centre=None):
    _chk(rows, BF16, "dlogits.rows"); _chk(cols, BF16, "dlogits.cols")
    m, e = rows.shape
    n = cols.shape[0]
    ld = (n + 7) // 8 * 8  # 16-byte row pitch for any batch size; columns [n, ld) are never written nor read
    out = torch.empty((m, ld), dtype=BF16, device=rows.device)
    with _profiled((m, n, e, L.EPI_CLIP_DLOGITS, False, False)):
        _call(L.lib().clipn_clip_dlogits(rows.data_ptr(), cols.data_ptr(), m, n, e, 1.0, scale.data_ptr(), label_offset,
                                           row_lse.data_ptr(), _ptr(col_lse), col_w, gscale, out.data_ptr(), ld,
                                           _ptr(scalar_acc), _ptr(row_centre), _stream()))
    return out


def clip_dfeat(dlogits, cols, alpha: torch.Tensor, n: Optional[int] = None, init: Optional[torch.Tensor] = None):
    """fp32 [m, E] = init + alpha * dlogits[:, :n] @ cols  (split-K over n, TMA reduce-add into `init` or zeros)."""
    _chk(dlogits, BF16, "dfeat.dlogits"); _chk(cols, BF16, "dfeat.cols")
    m, ld = dlogits.shape
    n = cols.shape[0] if n is None else n
    e = cols.shape[1]
    if init is not None:
        _chk(init, F32, "dfeat.init")
        assert tuple(init.shape) == (m, e)
    out = init if init is not None else torch.zeros((m, e), dtype=F32, device=dlogits.device)
    splits = wgrad_splits(m, e, n)
    with _profiled((m, e, n, L.EPI_ACCUM_F32, False, True)):
        _call(L.lib().clipn_clip_dfeat(dlogits.data_ptr(), ld, cols.data_ptr(), m, n, e, 1.0, alpha.data_ptr(),
                                         out.data_ptr(), splits, _stream()))
    return out


def siglip_dir(rows, cols, scale, bias, label_offset, gscale, loss_acc, scalar_acc, want_grad: bool):
    """Generic SigLIP direction on a LOCAL column operand (fallback for shapes outside the fused kernel's envelope):
    loss_acc / scalar_acc (optional) accumulate the value and d scale, d bias; returns d(logits) bf16 [m, n] or None."""
    m, e = rows.shape
    n = cols.shape[0]
    ld = (n + 7) // 8 * 8
    dl = torch.empty((m, ld), dtype=BF16, device=rows.device) if want_grad else None
    d = L.GemmDesc()
    d.a, d.lda, d.a_mn_major = rows.data_ptr(), e, 0
    d.b, d.ldb, d.b_mn_major = cols.data_ptr(), e, 0
    if dl is not None:
        d.c, d.ldc = dl.data_ptr(), ld
    d.m, d.n, d.k = m, n, e
    d.epilogue, d.alpha, d.splits = L.EPI_SIGLIP, 1.0, 1
    d.alpha_dev, d.logit_bias_dev = scale.data_ptr(), bias.data_ptr()
    d.gscale, d.label_offset, d.negative_only = gscale, label_offset, 0
    d.part_sum = _ptr(loss_acc)
    d.scalar_acc = _ptr(scalar_acc)
    _call(L.lib().clipn_gemm(C.byref(d), _stream()))
    return dl
