"""GPU parity: HBM-bound kernels vs plain torch fp32 references of the same op."""
import pytest
import torch
import torch.nn.functional as F

from open_clip_b200 import ops
from gpu_util import BF16, F32, max_err, randn, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,d", [(7, 128), (400, 768), (616, 512), (4096, 1024), (1, 64)])
def test_layernorm_fwd_bwd(rows, d):
    x = randn(rows, d, seed=1)
    gamma = (1 + 0.1 * randn(d, seed=2, dtype=F32)).contiguous()
    beta = 0.1 * randn(d, seed=3, dtype=F32)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (d,), gr, br, 1e-5)
    assert rel_err(y, yr) < 4e-3
    assert max_err(mean, xr.mean(-1)) < 1e-4
    dy = randn(rows, d, seed=4)
    resid = randn(rows, d, seed=5)
    dg = torch.zeros(d, dtype=F32, device="cuda")
    db = torch.zeros(d, dtype=F32, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db, resid=resid)
    yr.backward(dy.float())
    assert rel_err(dx, xr.grad + resid.float()) < 5e-3
    assert rel_err(dg, gr.grad) < 1e-3
    assert rel_err(db, br.grad) < 1e-3
    # fused bias gradient of the Linear that fed the residual stream: column sums of the residual gradient (+=)
    rs = torch.ones(d, dtype=F32, device="cuda")
    dg2, db2 = torch.zeros_like(dg), torch.zeros_like(db)
    dx2 = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg2, db2, resid=resid, resid_sum=rs)
    assert torch.equal(dx2, dx) and rel_err(dg2, dg) < 1e-5
    assert rel_err(rs, 1.0 + resid.float().sum(0)) < 1e-4


def test_patchify_matches_conv_unfold():
    B, P, H = 3, 16, 64
    img = randn(B, 3, H, H, seed=1)
    patches = ops.patchify(img, P)
    want = F.unfold(img.float(), kernel_size=P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)
    assert torch.equal(patches.float(), want)
    w = randn(128, 3, P, P, seed=2, scale=0.05)
    pe = ops.gemm(patches, w.view(128, -1))
    conv = F.conv2d(img.float(), w.float(), stride=P).reshape(B, 128, -1).permute(0, 2, 1).reshape(-1, 128)
    assert rel_err(pe, conv) < 4e-3


def test_patchify_patch14_padded_and_weight_grad_unpad():
    """ViT-L/14: 588 columns padded to 592 (zero-filled) for both the im2row matrix and the conv1 weight; the K-padded
    weight gradient is folded back into the [d, 588] parameter layout."""
    B, P, H, d = 2, 14, 56, 128
    img = randn(B, 3, H, H, seed=3)
    patches = ops.patchify(img, P)
    assert patches.shape == (B * 16, 592)
    want = F.unfold(img.float(), kernel_size=P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)
    assert torch.equal(patches[:, :588].float(), want) and float(patches[:, 588:].float().abs().max()) == 0.0
    w = randn(d, 3, P, P, seed=4, scale=0.05)
    wp = ops.patchify(w, P)
    assert torch.equal(wp[:, :588], w.view(d, -1)) and float(wp[:, 588:].float().abs().max()) == 0.0
    pe = ops.gemm(patches, wp)
    conv = F.conv2d(img.float(), w.float(), stride=P).reshape(B, d, -1).permute(0, 2, 1).reshape(-1, d)
    assert rel_err(pe, conv) < 4e-3
    gpad = torch.randn(d, 592, device="cuda")
    g = torch.randn(d, 588, device="cuda")
    want_g = g + gpad[:, :588]
    ops.accum_rows_f32(g, gpad, 588)
    assert torch.equal(g, want_g)


def test_vision_embed_fwd_bwd():
    B, npatch, d = 5, 16, 128
    pe = randn(B * npatch, d, seed=1)
    cls, pos = randn(d, seed=2, dtype=F32), randn(npatch + 1, d, seed=3, dtype=F32)
    x = ops.vision_embed_fwd(pe, cls, pos, B, npatch)
    want = torch.cat([cls.to(BF16).view(1, 1, d).expand(B, 1, d), pe.view(B, npatch, d)], 1) + pos.to(BF16)
    assert torch.equal(x.view(B, npatch + 1, d), want)
    dx = randn(B * (npatch + 1), d, seed=4)
    dcls, dpos = torch.zeros(d, device="cuda"), torch.zeros(npatch + 1, d, device="cuda")
    dpe = ops.vision_embed_bwd(dx, dcls, dpos, B, npatch)
    dx3 = dx.float().view(B, npatch + 1, d)
    assert torch.equal(dpe.view(B, npatch, d), dx.view(B, npatch + 1, d)[:, 1:])
    assert rel_err(dcls, dx3[:, 0].sum(0)) < 1e-5
    assert rel_err(dpos, dx3.sum(0)) < 1e-5


def test_text_embed_fwd_bwd():
    B, S, d, V = 6, 20, 128, 512
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1, V - 1, (B, S), generator=g)
    ids[:, -3] = V - 1
    ids = ids.cuda()
    table, pos = randn(V, d, seed=1, dtype=F32), randn(S, d, seed=2, dtype=F32)
    x, eot = ops.text_embed_fwd(ids, table, pos)
    want = F.embedding(ids, table).to(BF16) + pos.to(BF16)
    assert torch.equal(x.view(B, S, d), want)
    assert torch.equal(eot.long(), ids.argmax(-1))
    dx = randn(B * S, d, seed=3)
    dtab, dpos = torch.zeros(V, d, device="cuda"), torch.zeros(S, d, device="cuda")
    ops.text_embed_bwd(ids, dx, dtab, dpos)
    want_tab = torch.zeros(V, d, device="cuda").index_add_(0, ids.view(-1), dx.float())
    assert rel_err(dtab, want_tab) < 1e-5
    assert rel_err(dpos, dx.float().view(B, S, d).sum(0)) < 1e-5


def test_gather_scatter_rows():
    B, S, d = 4, 10, 64
    x = randn(B * S, d, seed=1)
    idx = torch.tensor([0, 9, 3, 5], dtype=torch.int32, device="cuda")
    out = ops.gather_rows(x, idx, B, S)
    assert torch.equal(out, x.view(B, S, d)[torch.arange(B), idx.long()])
    assert torch.equal(ops.gather_rows(x, None, B, S), x.view(B, S, d)[:, 0])
    dx = ops.scatter_rows(out, idx, B, S)
    want = torch.zeros(B, S, d, dtype=BF16, device="cuda")
    want[torch.arange(B), idx.long()] = out
    assert torch.equal(dx.view(B, S, d), want)


@pytest.mark.parametrize("dy_f32", [True, False])
def test_l2norm(dy_f32):
    x = randn(37, 512, seed=1)
    y, inv = ops.l2norm_fwd(x)
    xr = x.float().requires_grad_(True)
    yr = F.normalize(xr, dim=-1)
    assert rel_err(y, yr) < 4e-3
    dy = randn(37, 512, seed=2, dtype=F32 if dy_f32 else BF16)
    dx = ops.l2norm_bwd(dy, y, inv)
    yr.backward(dy.float())
    assert rel_err(dx, xr.grad) < 1e-2


def test_colsum_and_cast():
    x = randn(1000, 2304, seed=1)
    out = torch.zeros(2304, device="cuda")
    ops.colsum(x, out)
    assert rel_err(out, x.float().sum(0)) < 1e-5
    f = randn(1000, 7, seed=2, dtype=F32)
    assert torch.equal(ops.cast_f32_to_bf16(f), f.to(BF16))
