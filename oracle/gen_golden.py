"""Pin the oracle against the REAL reference and mint the fixtures in tests/golden/.

Run in the authoring container only (needs /root/reference, which does not exist on the
GPU box):

    python oracle/gen_golden.py

What it does
  1. imports open_clip from /root/reference/src (with a 2-line `ftfy` stub, the only
     missing hard dependency: src/open_clip/tokenizer.py:14);
  2. for `tiny`, `ViT-B-32`, the ViT-L-14-336 geometry at depth 2+2 (ClipLoss) and ViT-B-16 with
     the reference SigLipLoss, fp32 and --precision bf16: loads oracle.init_params() into
     the reference CLIP, runs reference forward + reference loss + backward, and
     asserts the oracle restatement reproduces features / loss / every parameter gradient;
  3. runs the reference ClipLoss / SigLipLoss under a real gloo process group (W=2,4) and
     asserts (W=2,4,8) the process-group-free restatement in clip_oracle reproduces values and
     feature gradients;
  4. writes the reference outputs as fixtures (features, losses, per-parameter gradient
     probes) that tests/ compare the oracle (CPU suite) and the CUDA path (GPU suite) to.
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"


def _import_reference():
    stub = tempfile.mkdtemp(prefix="ftfy_stub_")
    with open(os.path.join(stub, "ftfy.py"), "w") as f:
        f.write("def fix_text(s):\n    return s\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF_SRC)
    import open_clip  # noqa
    return open_clip


def _ref_model(open_clip, cfg, params, precision):
    from open_clip.model import CLIP, convert_weights_to_lp, get_cast_dtype
    vision_cfg = dict(image_size=cfg.image_size, layers=cfg.v_layers, width=cfg.v_width, patch_size=cfg.patch_size)
    text_cfg = dict(context_length=cfg.t_ctx, vocab_size=cfg.t_vocab, width=cfg.t_width, heads=cfg.t_heads,
                    layers=cfg.t_layers)
    kw = {}
    if "logit_bias" in params:
        kw["init_logit_bias"] = float(params["logit_bias"])
    model = CLIP(cfg.embed_dim, vision_cfg, text_cfg, cast_dtype=get_cast_dtype(precision), output_dict=True, **kw)
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if precision == "bf16":
        convert_weights_to_lp(model, dtype=torch.bfloat16)  # factory.py:889-918 path for 'bf16'
    return model


def grad_probe(name: str, g: torch.Tensor):
    """(L2 norm, projection on a fixed seeded direction) — a 2-number fingerprint per gradient."""
    gen = torch.Generator().manual_seed(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    d = torch.randn(g.shape, generator=gen)
    g = g.float()
    return torch.tensor([g.norm().item(), (g * d).sum().item()])


def model_goldens(open_clip, cfg_name, batch, seed, precisions=("fp32", "bf16"), keep_full_grads=False,
                  siglip=False):
    """siglip=True: reference SigLipLoss (world 1) on a model built with init_logit_scale=log(10),
    init_logit_bias=-10 (the reference's SigLIP defaults, model.py:326 / factory --siglip)."""
    from oracle import clip_oracle as O
    from open_clip.loss import ClipLoss, SigLipLoss
    cfg = O.CONFIGS[cfg_name]
    init_kw = dict(init_logit_scale=2.302585, init_logit_bias=-10.0) if siglip else {}
    base = O.init_params(cfg, seed=seed, bias_std=0.02, **init_kw)
    image, text = O.synthetic_batch(cfg, batch, seed=100 + seed)
    out = {"cfg": cfg_name, "batch": batch, "seed": seed, "image": image if cfg_name == "tiny" else None,
           "text": text, "param_checksum": float(sum(v.double().abs().sum() for v in base.values())),
           "image_checksum": float(image.double().abs().sum()), "siglip": siglip, "init_kw": init_kw}
    for prec in precisions:
        model = _ref_model(open_clip, cfg, base, prec)
        in_dtype = torch.bfloat16 if prec == "bf16" else torch.float32
        ref = model(image=image.to(in_dtype), text=text)
        if siglip:
            loss = SigLipLoss()(ref["image_features"], ref["text_features"], ref["logit_scale"], ref["logit_bias"])
        else:
            loss = ClipLoss()(ref["image_features"], ref["text_features"], ref["logit_scale"])
        loss.backward()
        ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}

        # ---- pin the oracle restatement to the reference
        p = {k: v.requires_grad_(True) for k, v in O.cast_params(base, prec).items()}
        o = O.clip_forward(p, cfg, image.to(in_dtype), text)
        if siglip:
            oloss = O.siglip_block_loss(o["image_features"], o["text_features"], o["logit_scale"], o["logit_bias"])
        else:
            oloss = O.clip_loss(o["image_features"], o["text_features"], o["logit_scale"])
        oloss.backward()
        tol = 0.0 if prec == "fp32" else 0.0
        for key in ("image_features", "text_features"):
            d = (o[key].float() - ref[key].float()).abs().max().item()
            assert d <= 1e-6 if prec == "fp32" else d <= 1e-2, (cfg_name, prec, key, d)
            print(f"  [{cfg_name}/{prec}] oracle vs reference {key}: max|d|={d:.3e}")
        dl = abs(float(oloss) - float(loss))
        print(f"  [{cfg_name}/{prec}] loss ref={float(loss):.6f} oracle={float(oloss):.6f}")
        assert dl <= (1e-5 if prec == "fp32" else 2e-2)
        worst = 0.0
        for k, g in ref_grads.items():
            og = p[k].grad
            rel = (og.float() - g.float()).norm().item() / (g.float().norm().item() + 1e-12)
            worst = max(worst, rel)
        print(f"  [{cfg_name}/{prec}] worst param-grad rel-L2 oracle vs reference: {worst:.3e}")
        assert worst <= (1e-4 if prec == "fp32" else 5e-2), worst

        out[prec] = {
            "image_features": ref["image_features"].detach().float().clone(),
            "text_features": ref["text_features"].detach().float().clone(),
            "loss": float(loss),
            "grad_probes": {k: grad_probe(k, g) for k, g in ref_grads.items()},
        }
        if keep_full_grads:
            out[prec]["grads"] = {k: g.float() for k, g in ref_grads.items()}
    return out


# ---------------------------------------------------------------- multi-rank loss goldens
def _loss_worker(rank, world, port, feats, kind, kwargs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _import_reference()
    from open_clip.loss import ClipLoss, SigLipLoss
    img = feats["img"][rank].clone().requires_grad_(True)
    txt = feats["txt"][rank].clone().requires_grad_(True)
    scale = feats["scale"].clone().requires_grad_(True)
    if kind == "clip":
        loss = ClipLoss(rank=rank, world_size=world, **kwargs)(img, txt, scale)
    else:
        bias = feats["bias"].clone().requires_grad_(True)
        loss = SigLipLoss(rank=rank, world_size=world, **kwargs)(img, txt, scale, bias)
    loss.backward()
    res = {"loss": float(loss), "d_img": img.grad.numpy().copy(), "d_txt": txt.grad.numpy().copy(),
           "d_scale": float(scale.grad)}
    if kind != "clip":
        res["d_bias"] = float(bias.grad)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def loss_goldens(world, B=8, E=64, seed=7, port=29611):
    from oracle import clip_oracle as O
    g = torch.Generator().manual_seed(seed + world)
    feats = {
        "img": [torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1) for _ in range(world)],
        "txt": [torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1) for _ in range(world)],
        "scale": torch.tensor(14.285714), "bias": torch.tensor(-10.0),
    }
    cases = [("clip", dict(local_loss=False, gather_with_grad=False)),
             ("clip", dict(local_loss=True, gather_with_grad=True)),
             ("clip", dict(local_loss=True, gather_with_grad=False)),
             ("clip", dict(local_loss=False, gather_with_grad=True)),
             ("siglip", dict(dist_impl="bidir")), ("siglip", dict(dist_impl="gather"))]
    out = {"world": world, "feats": feats, "cases": []}
    ctx = mp.get_context("spawn")
    for ci, (kind, kw) in enumerate(cases):
        q = ctx.Queue()
        procs = [ctx.Process(target=_loss_worker, args=(r, world, port + ci + 10 * world, feats, kind, kw, q))
                 for r in range(world)]
        [p.start() for p in procs]
        res = dict(q.get() for _ in range(world))
        [p.join() for p in procs]
        res = [res[r] for r in range(world)]
        for x in res:
            x["d_img"], x["d_txt"] = torch.from_numpy(x["d_img"]), torch.from_numpy(x["d_txt"])
        # ---- pin the process-group-free restatement
        img = [f.clone().requires_grad_(True) for f in feats["img"]]
        txt = [f.clone().requires_grad_(True) for f in feats["txt"]]
        scale = feats["scale"].clone().requires_grad_(True)
        if kind == "clip":
            losses = O.clip_loss_ranks(img, txt, scale, kw["local_loss"], kw["gather_with_grad"])
        else:
            bias = feats["bias"].clone().requires_grad_(True)
            losses = O.siglip_loss_ranks(img, txt, scale, bias)
        # every rank backprops its own loss; feature grads SUM over ranks' losses (what DDP sees pre-mean)
        sum(losses).backward()
        for r in range(world):
            assert abs(float(losses[r]) - res[r]["loss"]) < 1e-5, (kind, kw, r)
            di = (img[r].grad - res[r]["d_img"]).abs().max().item()
            dt = (txt[r].grad - res[r]["d_txt"]).abs().max().item()
            assert di < 1e-6 and dt < 1e-6, (kind, kw, r, di, dt)
        print(f"  [W={world}] {kind} {kw}: losses {[round(x['loss'], 6) for x in res]} restatement OK")
        out["cases"].append({"kind": kind, "kwargs": kw, "ranks": res})
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    open_clip = _import_reference()
    gold_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold_dir, exist_ok=True)
    only = sys.argv[1:] or ["tiny", "vitb32", "vitl14", "vitl14_full", "vitb16_siglip", "loss"]
    if "tiny" in only:
        print("model goldens: tiny")
        torch.save(model_goldens(open_clip, "tiny", batch=8, seed=1, keep_full_grads=False),
                   os.path.join(gold_dir, "tiny_model.pt"))
    if "vitb32" in only:
        print("model goldens: ViT-B-32")
        torch.save(model_goldens(open_clip, "ViT-B-32", batch=8, seed=0), os.path.join(gold_dir, "vitb32_model.pt"))
    if "vitl14" in only:
        print("model goldens: ViT-L-14-336 geometry, depth 2+2")
        torch.save(model_goldens(open_clip, "ViT-L-14-336-d2", batch=8, seed=11),
                   os.path.join(gold_dir, "vitl14_336_d2_model.pt"))
    if "vitl14_full" in only:
        # BASELINE config 4 at its real depth (24 + 12 blocks), batch 8: ~15 min of reference autograd on 8 CPU cores
        print("model goldens: ViT-L-14-336, full depth")
        torch.save(model_goldens(open_clip, "ViT-L-14-336", batch=8, seed=13),
                   os.path.join(gold_dir, "vitl14_336_full_model.pt"))
    if "vitb16_siglip" in only:
        print("model goldens: ViT-B-16 + SigLipLoss")
        torch.save(model_goldens(open_clip, "ViT-B-16", batch=8, seed=3, siglip=True),
                   os.path.join(gold_dir, "vitb16_siglip_model.pt"))
    for w in ((2, 4, 8) if "loss" in only else ((8,) if "loss8" in only else ())):
        print(f"loss goldens: world={w}")
        torch.save(loss_goldens(w), os.path.join(gold_dir, f"loss_w{w}.pt"))
    print("done")


if __name__ == "__main__":
    main()
