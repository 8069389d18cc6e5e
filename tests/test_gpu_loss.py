"""GPU parity: the fused contrastive-loss kernels (through NativeClipLoss / NativeSigLipLoss) vs the oracle's
restatement of the reference losses on identical bf16-representable inputs.  Tolerances: value 2e-3 abs,
feature grads rel-L2 1e-2 (d(logits) is rounded to bf16 once), logit-scale grad 1e-2 rel."""
import pytest
import torch
import torch.nn.functional as F

from open_clip_b200.loss import NativeClipLoss, NativeSigLipLoss
from oracle import clip_oracle as O
from gpu_util import BF16, F32, randn, rel_err

pytestmark = pytest.mark.gpu


def _feats(b, e, seed):
    i = F.normalize(randn(b, e, seed=seed, dtype=F32), dim=-1).to(BF16)
    t = F.normalize(randn(b, e, seed=seed + 1, dtype=F32), dim=-1).to(BF16)
    return i, t


@pytest.fixture(params=["fused", "generic"])
def path(request, monkeypatch):
    """fused = the peer-streaming kernel (stationary column tile; what multi-rank runs use, here with one rank);
    generic = the fallback for shapes outside its envelope (M-major tcgen05 GEMM with the same epilogues)."""
    if request.param == "generic":
        from open_clip_b200 import ops
        monkeypatch.setattr(ops, "peer_gemm_tile_n", lambda world, b, e: 0)
    return request.param


@pytest.mark.parametrize("b,e", [(8, 64), (32, 64), (128, 128), (200, 512), (1000, 512), (4096, 512), (520, 768),
                                 (256, 1024), (64, 32), (2, 768), (6, 64), (50, 512), (131, 64)])  # any batch size
@pytest.mark.parametrize("feat_dtype", [BF16, F32])
def test_clip_loss_value_and_grads(b, e, feat_dtype, path):
    i, t = _feats(b, e, 3)
    scale = torch.tensor(14.2857, device="cuda")
    gi, gt = i.clone().to(feat_dtype).requires_grad_(True), t.clone().to(feat_dtype).requires_grad_(True)
    gs = scale.clone().requires_grad_(True)
    out = NativeClipLoss()(gi, gt, gs, output_dict=True)
    loss = out["contrastive_loss"]
    loss.backward()
    ri, rt = i.detach().float().cpu().requires_grad_(True), t.detach().float().cpu().requires_grad_(True)
    rs = scale.cpu().clone().requires_grad_(True)
    rl = O.clip_loss(ri, rt, rs)
    rl.backward()
    assert abs(float(loss) - float(rl)) < (2e-3 if feat_dtype == F32 else 3e-2)
    assert rel_err(gi.grad.cpu(), ri.grad) < 1e-2
    assert rel_err(gt.grad.cpu(), rt.grad) < 1e-2
    assert abs(float(gs.grad) - float(rs.grad)) < 1e-2 * abs(float(rs.grad)) + 1e-5


def test_clip_loss_large_logit_scale():
    """logit_scale at its clamp (100, image_text_task.py:98-101): the online LSE must not over/underflow."""
    b, e = 512, 512
    i, t = _feats(b, e, 9)
    t = F.normalize(t.float() + 0.1 * i.float(), dim=-1).to(BF16)  # positives ~10 logits above the negatives' spread
    scale = torch.tensor(100.0, device="cuda")
    gi, gt, gs = i.clone().requires_grad_(True), t.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    loss = NativeClipLoss()(gi, gt, gs)
    loss.backward()
    ri, rt = i.float().cpu().requires_grad_(True), t.float().cpu().requires_grad_(True)
    rs = scale.cpu().clone().requires_grad_(True)
    rl = O.clip_loss(ri, rt, rs)
    rl.backward()
    assert bool(torch.isfinite(loss)), float(loss)
    assert abs(float(loss) - float(rl)) < 3e-2 + 1e-2 * abs(float(rl)), (float(loss), float(rl))
    assert rel_err(gi.grad.cpu(), ri.grad) < 1.5e-2, rel_err(gi.grad.cpu(), ri.grad)
    assert rel_err(gt.grad.cpu(), rt.grad) < 1.5e-2


def test_clip_loss_gradients_of_nearly_parallel_features():
    """Features of a freshly initialised model are almost identical across the batch (cos ~ 0.999): P - onehot then
    multiplies nearly equal vectors and the result is a small difference of O(1) terms.  The one-hot part is kept out
    of the bf16 d(logits) tiles for exactly this case (this is the shape of bench.py's parity block)."""
    b, e = 1024, 512
    g = torch.Generator().manual_seed(2)
    mu = F.normalize(torch.randn(1, e, generator=g), dim=-1)
    i = F.normalize(mu + 0.02 * torch.randn(b, e, generator=g), dim=-1).to(BF16).cuda()
    t = F.normalize(mu + 0.02 * torch.randn(b, e, generator=g), dim=-1).to(BF16).cuda()
    scale = torch.tensor(14.2857, device="cuda")
    gi, gt, gs = i.clone().requires_grad_(True), t.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    loss = NativeClipLoss()(gi, gt, gs)
    loss.backward()
    ri, rt = i.float().cpu().requires_grad_(True), t.float().cpu().requires_grad_(True)
    rs = scale.cpu().clone().requires_grad_(True)
    rl = O.clip_loss(ri, rt, rs)
    rl.backward()
    assert abs(float(loss) - float(rl)) < 1e-2
    assert rel_err(gi.grad.cpu(), ri.grad) < 1.5e-2, rel_err(gi.grad.cpu(), ri.grad)
    assert rel_err(gt.grad.cpu(), rt.grad) < 1.5e-2, rel_err(gt.grad.cpu(), rt.grad)


def test_siglip_no_grad_forward_skips_the_gradient_gemms():
    from open_clip_b200 import ops
    i, t = _feats(256, 512, 5)
    scale, bias = torch.tensor(10.0, device="cuda"), torch.tensor(-10.0, device="cuda")
    ops.LAUNCHES = 0
    with torch.no_grad():
        l0 = NativeSigLipLoss()(i.clone().requires_grad_(True), t, scale, bias)
    assert ops.LAUNCHES == 1  # one fused launch: no d(logits), no d(feature) GEMMs
    l1 = NativeSigLipLoss()(i.clone().requires_grad_(True), t, scale, bias)
    assert abs(float(l0) - float(l1)) < 1e-3 * abs(float(l1)) + 1e-4


@pytest.mark.parametrize("b,e", [(64, 32), (256, 512), (200, 128), (1000, 512), (6, 64), (131, 128)])
def test_siglip_loss_value_and_grads(b, e, path):
    i, t = _feats(b, e, 5)
    scale, bias = torch.tensor(10.0, device="cuda"), torch.tensor(-10.0, device="cuda")
    gi, gt = i.clone().requires_grad_(True), t.clone().requires_grad_(True)
    gs, gb = scale.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    loss = NativeSigLipLoss()(gi, gt, gs, gb)
    loss.backward()
    ri, rt = i.float().cpu().requires_grad_(True), t.float().cpu().requires_grad_(True)
    rs, rb = scale.cpu().clone().requires_grad_(True), bias.cpu().clone().requires_grad_(True)
    rl = O.siglip_block_loss(ri, rt, rs, rb)
    rl.backward()
    assert abs(float(loss) - float(rl)) < 2e-2 * abs(float(rl)) + 1e-3
    assert rel_err(gi.grad.cpu(), ri.grad) < 1.5e-2
    assert rel_err(gt.grad.cpu(), rt.grad) < 1.5e-2
    assert abs(float(gs.grad) - float(rs.grad)) < 2e-2 * abs(float(rs.grad)) + 1e-4
    assert abs(float(gb.grad) - float(rb.grad)) < 2e-2 * abs(float(rb.grad)) + 1e-4
