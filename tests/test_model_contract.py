"""CPU suite: the drop-in boundary of NativeCLIP that does not need a kernel launch — parameter
names/shapes/dtypes, the attributes and methods the reference's task/optimizer/main code touches, and the
"no CPU fallback" rule. Where /root/reference is present (authoring container only) the parameter table is also
compared with the real reference CLIP."""
import os
import sys

import pytest
import torch

from open_clip_b200._lib import ClipnError
from open_clip_b200.loss import NativeClipLoss, NativeSigLipLoss
from open_clip_b200.model import CONFIGS, NativeCLIP, create_model
from oracle import clip_oracle as O

BF16, F32 = torch.bfloat16, torch.float32


def _small(name):
    """The named geometry with one block per tower (parameter contract only; keeps the CPU suite fast)."""
    c = CONFIGS[name]
    v, t = dict(c["vision_cfg"], layers=1), dict(c["text_cfg"], layers=1)
    return NativeCLIP(c["embed_dim"], v, t, output_dict=True, device="cpu"), O.ClipCfg(
        embed_dim=c["embed_dim"], image_size=v["image_size"], patch_size=v["patch_size"], v_width=v["width"], v_layers=1,
        t_ctx=t["context_length"], t_vocab=t["vocab_size"], t_width=t["width"], t_heads=t["heads"], t_layers=1)


@pytest.mark.parametrize("name", ["ViT-B-32", "ViT-B-16", "ViT-L-14-336", "tiny"])
def test_parameter_names_shapes_dtypes(name):
    m, cfg = _small(name)
    shapes = O.param_shapes(cfg)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
        assert v.dtype == (BF16 if O.is_lowp_param(k) else F32), k  # convert_weights_to_lp contract (model.py:738-765)
    assert all(isinstance(p, torch.nn.Parameter) for p in m.parameters())
    assert m.visual.image_size == (cfg.image_size, cfg.image_size)
    assert m.context_length == cfg.t_ctx and m.vocab_size == cfg.t_vocab
    assert isinstance(m.logit_scale, torch.nn.Parameter) and m.logit_scale.ndim == 0
    assert "attn_mask" not in sd  # non-persistent buffer, as in the reference (model.py:360)


def test_full_vitb32_parameter_count_and_optimizer_grouping_inputs():
    m = create_model("ViT-B-32", device="cpu")
    assert sum(p.numel() for p in m.parameters()) == 151277313
    assert m.no_weight_decay() == {"positional_embedding", "visual.positional_embedding", "visual.class_embedding"}
    # optim.py:67-75 splits weight decay by ndim: every gain/bias must stay <= 1-D, every matrix >= 2-D
    for n, p in m.named_parameters():
        if n.endswith(".bias") or ".ln_" in n or n.startswith(("ln_final", "visual.ln_")) or n == "logit_scale":
            assert p.ndim <= 1, n
        if n.endswith("in_proj_weight") or n.endswith("c_fc.weight") or n == "visual.conv1.weight":
            assert p.ndim >= 2, n


def test_no_cpu_fallback_and_loss_requires_cuda():
    m, cfg = _small("tiny")
    image, text = O.synthetic_batch(cfg, 2, seed=0)
    with pytest.raises(ClipnError):
        m(image=image.to(BF16), text=text)
    with pytest.raises(ClipnError):
        m.encode_text(text)
    f = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1).to(BF16)
    with pytest.raises((ClipnError, RuntimeError, AssertionError)):
        NativeClipLoss()(f, f, torch.tensor(10.0))
    with pytest.raises((ClipnError, RuntimeError, AssertionError)):
        NativeSigLipLoss()(f, f, torch.tensor(10.0), torch.tensor(-10.0))


def test_lock_towers_follow_reference_group_semantics():
    """lock_image_tower / lock_text_tower (main.py:315-326): unlocked_* counts the TOP groups left trainable, proj
    first, the last block travels with ln_post / ln_final; every call sets all groups explicitly."""
    c = CONFIGS["tiny"]
    m = NativeCLIP(c["embed_dim"], c["vision_cfg"], c["text_cfg"], device="cpu")
    names = [n for n, _ in m.visual.layer_groups()]
    assert names == ["embeddings", "layer.0", "layer.1", "proj"]
    m.lock_image_tower(unlocked_groups=0)
    assert not any(p.requires_grad for n, p in m.named_parameters() if n.startswith("visual."))
    assert all(p.requires_grad for n, p in m.named_parameters() if not n.startswith("visual."))
    m.lock_image_tower(unlocked_groups=2)  # proj + (last block, ln_post)
    on = {n for n, p in m.named_parameters() if n.startswith("visual.") and p.requires_grad}
    assert "visual.proj" in on and "visual.ln_post.weight" in on
    assert any(n.startswith("visual.transformer.resblocks.1.") for n in on)
    assert not any(n.startswith("visual.transformer.resblocks.0.") for n in on)
    assert "visual.conv1.weight" not in on and "visual.ln_pre.weight" not in on
    m.lock_text_tower(unlocked_layers=1)  # only text_projection stays trainable
    text_on = {n for n, p in m.named_parameters() if not n.startswith("visual.") and p.requires_grad}
    assert text_on == {"text_projection", "logit_scale"}
    m.lock_text_tower(unlocked_layers=0)
    assert not m.text_projection.requires_grad and not m.ln_final.weight.requires_grad
    assert [n for n, _ in m.text_layer_groups()] == ["embeddings", "layer.0", "layer.1", "proj"]


def test_state_dict_round_trip_from_reference_layout_and_siglip_bias():
    c = CONFIGS["tiny"]
    cfg = O.CONFIGS["tiny"]
    base = O.init_params(cfg, seed=5, bias_std=0.02, init_logit_scale=2.302585, init_logit_bias=-10.0)
    m = NativeCLIP(c["embed_dim"], c["vision_cfg"], c["text_cfg"], init_logit_bias=-10.0, device="cpu")
    m.load_reference_state_dict(base)
    for k, v in m.state_dict().items():
        want = base[k].to(v.dtype)
        assert torch.equal(v, want), k
    assert "logit_bias" in m.state_dict() and float(m.logit_bias) == -10.0
    with pytest.raises(ClipnError):
        m.load_reference_state_dict({k: v for k, v in base.items() if k != "ln_final.weight"})
    m.set_grad_checkpointing(True)
    assert m.grad_checkpointing is True


def test_returned_grads_never_alias_the_gradient_arena(monkeypatch):
    """The towers accumulate weight gradients in a flat fp32 arena that every backward zeroes and refills, so what
    autograd receives must be storage owned by that call: (a) `--accum-freq` (train.py:236-311) calls backward several
    times before optimizer.step(); (b) one loss may use the same tower twice in ONE graph (multi-view / multi-caption):
    the first node's gradients are still in the engine's input buffers when the second node runs."""
    from open_clip_b200 import ops, tower
    from open_clip_b200.model import _TowerFn
    c = CONFIGS["tiny"]
    m = NativeCLIP(c["embed_dim"], c["vision_cfg"], c["text_cfg"], device="cpu")
    calls = {"n": 0}

    def fake_fwd(P, cfg, inp, normalize, ws, save, checkpoint=False):
        return torch.zeros(inp.shape[0], cfg.embed_dim), (tower.TowerSaved(batch=inp.shape[0]) if save else None)

    def fake_bwd(P, G, cfg, saved, dfeat, ws):
        calls["n"] += 1
        for g in G.values():
            g.add_(float(10 ** (calls["n"] - 1)))  # 1 for the first node that runs, 10 for the second

    monkeypatch.setattr(tower, "vision_forward", fake_fwd)
    monkeypatch.setattr(tower, "vision_backward", fake_bwd)
    monkeypatch.setattr(ops, "cast_f32_to_bf16", lambda x, out=None: out.copy_(x.to(BF16)))
    params = dict(m.named_parameters())
    plist = [params[n] for n in m._tower_param_names["visual"]]
    image = torch.zeros(2, 3, 64, 64)
    two_views = _TowerFn.apply(m, "visual", True, True, image, *plist) + _TowerFn.apply(m, "visual", True, True, image, *plist)
    two_views.sum().backward()
    assert calls["n"] == 2
    arena = m._grad_arena("visual")
    for n, p in zip(m._tower_param_names["visual"], plist):
        assert float(p.grad.float().mean()) == 11.0, n  # grad_first + grad_second, not 2 * grad_second
        assert p.grad.dtype == p.dtype
        assert p.grad.data_ptr() != arena["views32"][n].data_ptr()
    arena["flat32"].zero_()
    assert all(float(p.grad.float().mean()) == 11.0 for p in plist)


def test_tower_autograd_glue_accumulates_across_backward_calls(monkeypatch):
    """The autograd node of a tower (one Function per tower, every parameter an input) with the kernel schedules
    replaced by recorders: gradients come back with each parameter's dtype, a second backward without zero_grad adds
    to the first (micro-batch accumulation), and zero_grad(set_to_none=True) starts over."""
    from open_clip_b200 import ops, tower
    from open_clip_b200.model import _TowerFn
    c = CONFIGS["tiny"]
    m = NativeCLIP(c["embed_dim"], c["vision_cfg"], c["text_cfg"], device="cpu")
    calls = {"n": 0}

    def fake_fwd(P, cfg, inp, normalize, ws, save, checkpoint=False):
        return torch.zeros(inp.shape[0], cfg.embed_dim), (tower.TowerSaved(batch=inp.shape[0]) if save else None)

    def fake_bwd(P, G, cfg, saved, dfeat, ws):
        calls["n"] += 1
        for g in G.values():
            g.add_(float(calls["n"]))  # the kernels accumulate into the arena the node zeroed

    monkeypatch.setattr(tower, "vision_forward", fake_fwd)
    monkeypatch.setattr(tower, "vision_backward", fake_bwd)
    monkeypatch.setattr(ops, "cast_f32_to_bf16", lambda x, out=None: out.copy_(x.to(BF16)))
    params = dict(m.named_parameters())
    plist = [params[n] for n in m._tower_param_names["visual"]]
    image = torch.zeros(2, 3, 64, 64)

    def run():
        _TowerFn.apply(m, "visual", True, True, image, *plist).sum().backward()

    run()
    for p in plist:
        assert p.grad is not None and p.grad.dtype == p.dtype and float(p.grad.float().mean()) == 1.0
    run()  # no zero_grad in between: 1 + 2
    assert all(float(p.grad.float().mean()) == 3.0 for p in plist)
    for p in plist:
        p.grad = None
    run()
    assert all(float(p.grad.float().mean()) == 3.0 for p in plist)
    assert all(p.grad is None for n, p in params.items() if not n.startswith("visual."))
    # a frozen parameter gets no gradient and does not disturb the others (lock_image_tower path)
    m.lock_image_tower(unlocked_groups=1)
    for p in plist:
        p.grad = None
    run()
    assert params["visual.proj"].grad is not None and params["visual.conv1.weight"].grad is None


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="authoring container only")
def test_reference_task_optimizer_and_train_step_drive_the_native_objects(monkeypatch):
    """The reference's OWN CLIPTask, create_optimizer and train-step closure (clip_task.py:28-46, optim.py:336,
    train.py:163-185, image_text_task.py:91-101) operate on NativeCLIP + NativeClipLoss unchanged. The two kernel
    entry points (tower forward, loss forward) are replaced by the oracle's CPU math, so what is exercised is the
    plumbing: keyword names, dict keys, parameter discovery, weight-decay grouping, in-place logit_scale clamp."""
    import tempfile
    stub = tempfile.mkdtemp(prefix="ftfy_stub_")
    with open(os.path.join(stub, "ftfy.py"), "w") as f:
        f.write("def fix_text(s):\n    return s\n")
    sys.path[:0] = [stub, "/root/reference/src"]
    try:
        from contextlib import nullcontext
        from types import SimpleNamespace
        from open_clip.task.clip_task import CLIPTask
        from open_clip_train.optim import OptimizerCfg, create_optimizer
        from open_clip_train.train import _make_train_step_no_accum_no_scaler
        c, cfg = CONFIGS["tiny"], O.CONFIGS["tiny"]
        model = NativeCLIP(c["embed_dim"], c["vision_cfg"], c["text_cfg"], output_dict=True, device="cpu")

        def cpu_tower(self, which, inp, normalize):
            p = {k: v.float() for k, v in self.named_parameters()}
            enc = O.encode_image if which == "visual" else O.encode_text
            return enc(p, cfg, inp.float() if which == "visual" else inp, normalize=normalize)

        def cpu_loss(self, image_features, text_features, logit_scale, logit_bias=None, output_dict=False):
            loss = O.clip_loss(image_features, text_features, logit_scale)
            return {"contrastive_loss": loss} if output_dict else loss

        monkeypatch.setattr(NativeCLIP, "_run_tower", cpu_tower)
        monkeypatch.setattr(NativeClipLoss, "forward", cpu_loss)
        task = CLIPTask(model, loss=NativeClipLoss(), device=torch.device("cpu"), verbose=False)
        assert task.trainable_module is model and isinstance(task.loss, NativeClipLoss)
        opt = create_optimizer(task.trainable_module, OptimizerCfg(lr=1e-3, weight_decay=0.2, beta1=0.9, beta2=0.98, eps=1e-6))
        decayed = {id(p) for g in opt.param_groups if g["weight_decay"] > 0 for p in g["params"]}
        named = dict(model.named_parameters())
        assert id(named["visual.conv1.weight"]) in decayed and id(named["text_projection"]) in decayed
        for n in ("positional_embedding", "visual.positional_embedding", "visual.class_embedding", "logit_scale",
                  "ln_final.weight", "visual.transformer.resblocks.0.mlp.c_fc.bias"):
            assert id(named[n]) not in decayed, n
        assert sum(len(g["params"]) for g in opt.param_groups) == len(named)
        step = _make_train_step_no_accum_no_scaler(task, opt, nullcontext, SimpleNamespace(grad_clip_norm=1.0))
        image, text = O.synthetic_batch(cfg, 8, seed=3)
        before = {k: v.detach().clone() for k, v in named.items()}
        task.train()
        losses, report = step({"image": image, "text": text})
        assert set(losses) >= {"contrastive_loss", "loss"} and torch.isfinite(losses["loss"])
        assert "logit_scale" in report
        assert all(p.grad is not None for p in model.parameters())
        assert any(not torch.equal(before[k], v) for k, v in named.items())  # optimizer.step() updated our parameters
        with torch.no_grad():
            model.logit_scale.fill_(9.0)
        task.clamp_logit_scale()
        assert abs(float(model.logit_scale.detach()) - 4.605170185988092) < 1e-6  # ln(100), in place on our Parameter

        # SigLIPTask (siglip_task.py:8-45): same flow, model_out additionally carries logit_bias (model.py:540-541)
        from open_clip.task.siglip_task import SigLIPTask

        def cpu_siglip(self, image_features, text_features, logit_scale, logit_bias, output_dict=False):
            loss = O.siglip_block_loss(image_features, text_features, logit_scale, logit_bias)
            return {"contrastive_loss": loss} if output_dict else loss

        monkeypatch.setattr(NativeSigLipLoss, "forward", cpu_siglip)
        sig_model = NativeCLIP(c["embed_dim"], c["vision_cfg"], c["text_cfg"], init_logit_scale=2.302585,
                               init_logit_bias=-10.0, output_dict=True, device="cpu")
        sig_task = SigLIPTask(sig_model, loss=NativeSigLipLoss(), device=torch.device("cpu"), verbose=False)
        sig_opt = create_optimizer(sig_task.trainable_module, OptimizerCfg(lr=1e-3))
        sig_step = _make_train_step_no_accum_no_scaler(sig_task, sig_opt, nullcontext, SimpleNamespace())
        sig_task.train()
        losses, report = sig_step({"image": image, "text": text})
        assert torch.isfinite(losses["loss"]) and sig_model.logit_bias.grad is not None
        assert "logit_bias" in report and "logit_scale" in report
    finally:
        del sys.path[:2]
        for k in [k for k in sys.modules if k == "ftfy" or k.startswith(("open_clip.", "open_clip_train"))
                  or k == "open_clip"]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="authoring container only")
def test_parameter_table_equals_the_real_reference_clip():
    import tempfile
    stub = tempfile.mkdtemp(prefix="ftfy_stub_")
    with open(os.path.join(stub, "ftfy.py"), "w") as f:
        f.write("def fix_text(s):\n    return s\n")
    sys.path[:0] = [stub, "/root/reference/src"]
    try:
        from open_clip.model import CLIP, convert_weights_to_lp
        c = CONFIGS["ViT-B-32"]
        ref = CLIP(c["embed_dim"], dict(c["vision_cfg"], layers=2), dict(c["text_cfg"], layers=2), output_dict=True)
        convert_weights_to_lp(ref, dtype=torch.bfloat16)
        mine = NativeCLIP(c["embed_dim"], dict(c["vision_cfg"], layers=2), dict(c["text_cfg"], layers=2), device="cpu")
        rs, ms = ref.state_dict(), mine.state_dict()
        assert list(rs) == list(ms)  # same names in the same order
        for k in rs:
            assert rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype, k
        assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
        assert [n for n, _ in ref.visual.layer_groups()] == [n for n, _ in mine.visual.layer_groups()]
        # loss plug points: same constructor and forward parameters, same defaults (loss.py:59-66,118; :324-331,406)
        import inspect
        from open_clip.loss import ClipLoss, SigLipLoss
        for r, mn in ((ClipLoss, NativeClipLoss), (SigLipLoss, NativeSigLipLoss)):
            for fn in ("__init__", "forward"):
                pr = inspect.signature(getattr(r, fn)).parameters
                pm = inspect.signature(getattr(mn, fn)).parameters
                assert list(pr) == list(pm), (r.__name__, fn)
                assert [p.default for p in pr.values()] == [p.default for p in pm.values()], (r.__name__, fn)
    finally:
        del sys.path[:2]
        for k in [k for k in sys.modules if k == "ftfy" or k.startswith("open_clip")]:
            if k.startswith("open_clip_b200"):
                continue
            del sys.modules[k]
