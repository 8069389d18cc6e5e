"""CPU suite: the oracle against the committed golden fixtures (minted from the real reference by
oracle/gen_golden.py) and against the reference's own pinned identity (tests/test_siglip_chunked_loss.py)."""
import os

import pytest
import torch

from oracle import clip_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _probe(name, g):
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    d = torch.randn(g.shape, generator=gen)
    g = g.float()
    return torch.tensor([g.norm().item(), (g * d).sum().item()])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_tiny_model_matches_reference_fixture(golden_dir, prec):
    gold = _load(golden_dir, "tiny_model.pt")
    cfg = O.CONFIGS["tiny"]
    base = O.init_params(cfg, seed=gold["seed"], bias_std=0.02)
    assert abs(float(sum(v.double().abs().sum() for v in base.values())) - gold["param_checksum"]) < 1e-6 * gold["param_checksum"]
    p = {k: v.requires_grad_(True) for k, v in O.cast_params(base, prec).items()}
    image = gold["image"].to(torch.bfloat16 if prec == "bf16" else torch.float32)
    out = O.clip_forward(p, cfg, image, gold["text"])
    loss = O.clip_loss(out["image_features"], out["text_features"], out["logit_scale"])
    loss.backward()
    g = gold[prec]
    tol = 1e-6 if prec == "fp32" else 2e-2
    assert (out["image_features"].float() - g["image_features"]).abs().max() <= tol
    assert (out["text_features"].float() - g["text_features"]).abs().max() <= tol
    assert abs(float(loss) - g["loss"]) <= (1e-5 if prec == "fp32" else 2e-2)
    for k, pr in g["grad_probes"].items():
        mine = _probe(k, p[k].grad)
        assert abs(mine[0] - pr[0]) <= (1e-4 if prec == "fp32" else 5e-2) * (pr[0].abs() + 1e-6), k


def test_vitb32_fp32_features_match_reference_fixture(golden_dir):
    gold = _load(golden_dir, "vitb32_model.pt")
    cfg = O.CONFIGS["ViT-B-32"]
    base = O.init_params(cfg, seed=gold["seed"], bias_std=0.02)
    image, text = O.synthetic_batch(cfg, gold["batch"], seed=100 + gold["seed"])
    assert abs(float(image.double().abs().sum()) - gold["image_checksum"]) < 1e-6 * gold["image_checksum"]
    with torch.no_grad():
        out = O.clip_forward(base, cfg, image, text)
        loss = O.clip_loss(out["image_features"], out["text_features"], out["logit_scale"])
    g = gold["fp32"]
    assert (out["image_features"] - g["image_features"]).abs().max() <= 1e-5
    assert (out["text_features"] - g["text_features"]).abs().max() <= 1e-5
    assert abs(float(loss) - g["loss"]) <= 1e-5


@pytest.mark.parametrize("fixture,cfg_name", [("vitl14_336_d2_model.pt", "ViT-L-14-336-d2"),
                                              ("vitl14_336_full_model.pt", "ViT-L-14-336"),  # config 4 at 24 + 12 blocks
                                              ("vitb16_siglip_model.pt", "ViT-B-16")])
def test_other_baseline_geometries_match_reference_fixture(golden_dir, fixture, cfg_name):
    """BASELINE config 4 geometry (patch 14, 577 tokens, widths 1024/768; depth 2+2) with ClipLoss and config 5
    (ViT-B-16, 197 tokens) with the reference SigLipLoss: oracle vs the reference's fp32 outputs."""
    gold = _load(golden_dir, fixture)
    cfg = O.CONFIGS[cfg_name]
    base = O.init_params(cfg, seed=gold["seed"], bias_std=0.02, **gold["init_kw"])
    assert abs(float(sum(v.double().abs().sum() for v in base.values())) - gold["param_checksum"]) < 1e-6 * gold["param_checksum"]
    image, text = O.synthetic_batch(cfg, gold["batch"], seed=100 + gold["seed"])
    assert abs(float(image.double().abs().sum()) - gold["image_checksum"]) < 1e-6 * gold["image_checksum"]
    with torch.no_grad():
        out = O.clip_forward(base, cfg, image, text)
        if gold["siglip"]:
            loss = O.siglip_block_loss(out["image_features"], out["text_features"], out["logit_scale"], out["logit_bias"])
        else:
            loss = O.clip_loss(out["image_features"], out["text_features"], out["logit_scale"])
    g = gold["fp32"]
    assert (out["image_features"] - g["image_features"]).abs().max() <= 1e-5
    assert (out["text_features"] - g["text_features"]).abs().max() <= 1e-5
    assert abs(float(loss) - g["loss"]) <= 1e-5 * max(1.0, abs(g["loss"]))


def test_clip_loss_is_invariant_to_logit_bias():
    """ClipLoss with a logit_bias (loss.py:100-116 adds it to every logit): the value does not change and the bias
    gets a zero gradient — the identity NativeClipLoss relies on when a model built with init_logit_bias is trained
    with the softmax loss."""
    g = torch.Generator().manual_seed(0)
    img = torch.nn.functional.normalize(torch.randn(16, 32, generator=g), dim=-1).double().requires_grad_(True)
    txt = torch.nn.functional.normalize(torch.randn(16, 32, generator=g), dim=-1).double().requires_grad_(True)
    scale = torch.tensor(14.0, dtype=torch.float64)
    bias = torch.tensor(-10.0, dtype=torch.float64, requires_grad=True)
    with_bias = O.clip_loss(img, txt, scale, bias)
    without = O.clip_loss(img, txt, scale)
    assert abs(float(with_bias) - float(without)) < 1e-12
    gi, gb = torch.autograd.grad(with_bias, (img, bias))
    gi0, = torch.autograd.grad(without, (img,))
    assert abs(float(gb)) < 1e-12 and (gi - gi0).abs().max() < 1e-12


@pytest.mark.parametrize("world", [2, 4, 8])
def test_per_rank_gradient_view_matches_gloo_reference_fixture(golden_dir, world):
    """clip_loss_rank_grads / siglip_loss_rank_grads (the memory-lean per-rank view bench.py's parity block and the
    multi-GPU tests use) against the outputs of the REAL reference under gloo, every rank, every mode."""
    gold = _load(golden_dir, f"loss_w{world}.pt")
    f = gold["feats"]
    for case in gold["cases"]:
        for r in range(world):
            want = case["ranks"][r]
            if case["kind"] == "clip":
                kw = case["kwargs"]
                v, di, dt, ds = O.clip_loss_rank_grads(f["img"], f["txt"], f["scale"], r, kw["local_loss"],
                                                       kw["gather_with_grad"])
            else:
                v, di, dt, ds, db = O.siglip_loss_rank_grads(f["img"], f["txt"], f["scale"], f["bias"], r)
                assert abs(float(db) - want["d_bias"]) < 1e-4 * max(1.0, abs(want["d_bias"]))
            assert abs(float(v) - want["loss"]) < 1e-5
            assert (di - want["d_img"]).abs().max() < 1e-6 and (dt - want["d_txt"]).abs().max() < 1e-6
            assert abs(float(ds) - want["d_scale"]) < 1e-4 * max(1.0, abs(want["d_scale"]))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_losses_match_gloo_reference_fixture(golden_dir, world):
    gold = _load(golden_dir, f"loss_w{world}.pt")
    f = gold["feats"]
    for case in gold["cases"]:
        img = [t.clone().requires_grad_(True) for t in f["img"]]
        txt = [t.clone().requires_grad_(True) for t in f["txt"]]
        scale = f["scale"].clone().requires_grad_(True)
        if case["kind"] == "clip":
            losses = O.clip_loss_ranks(img, txt, scale, case["kwargs"]["local_loss"], case["kwargs"]["gather_with_grad"])
        else:
            losses = O.siglip_loss_ranks(img, txt, scale, f["bias"].clone().requires_grad_(True))
        sum(losses).backward()
        for r in range(world):
            assert abs(float(losses[r]) - case["ranks"][r]["loss"]) < 1e-5
            assert (img[r].grad - case["ranks"][r]["d_img"]).abs().max() < 1e-6
            assert (txt[r].grad - case["ranks"][r]["d_txt"]).abs().max() < 1e-6


def test_siglip_chunked_identity():
    """The one value identity the reference's own tests pin for this path (tests/test_siglip_chunked_loss.py):
    softplus form with a diagonal correction == labels form.  B=64, D=32, scale 10, bias -10, atol 1e-5."""
    g = torch.Generator().manual_seed(0)
    img = torch.nn.functional.normalize(torch.randn(64, 32, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(torch.randn(64, 32, generator=g), dim=-1)
    scale, bias = torch.tensor(10.0), torch.tensor(-10.0)
    full = O.siglip_block_loss(img, txt, scale, bias)
    logits = scale * img @ txt.T + bias
    chunked = (torch.nn.functional.softplus(logits).sum() - logits.diag().sum()) / 64
    assert abs(float(full) - float(chunked)) < 1e-5
    neg = O.siglip_block_loss(img, txt, scale, bias, negative_only=True)
    assert abs(float(neg) - float(torch.nn.functional.softplus(logits).sum() / 64)) < 1e-5


def test_cpu_trainer_steps_and_loss_decreases():
    tr = O.CpuTrainer(O.CONFIGS["tiny"], seed=0, lr=1e-3)
    image, text = O.synthetic_batch(O.CONFIGS["tiny"], 8, seed=3)
    losses = [tr.step(image, text) for _ in range(4)]
    assert losses[-1] < losses[0]
