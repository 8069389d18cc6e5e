#!/usr/bin/env python
"""bench.py — image-text pairs/s of one full CLIP train step (ViT-B-32, bf16, local batch 4096/GPU).

  python bench.py --gpus 1 --steps K --warmup W                      # this repo (sm_100a kernels)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --steps K --warmup W              # reference CPU train step (oracle port)

A "step" = H2D of the batch (e2e only) -> NativeCLIP.forward (both towers) -> NativeClipLoss -> backward ->
fused AdamW (reference ViT defaults) -> clamp logit_scale, i.e. exactly what the reference's
_make_train_step_no_accum_no_scaler + clamp_logit_scale do (open_clip_train/train.py:163-185,406).
Prints ONE JSON line on rank 0 (contract: see the task statement / DESIGN.md §measurement).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FWD_GFLOP_PER_PAIR = 14.78          # docs/model_profile.csv:8 (ViT-B-32), SURVEY §8d
STEP_GFLOP_PER_PAIR = 3 * FWD_GFLOP_PER_PAIR
# forward GFLOP per pair of the other BASELINE models (docs/model_profile.csv); step = 3x (no recompute)
FWD_GFLOP_BY_MODEL = {"ViT-B-32": FWD_GFLOP_PER_PAIR, "ViT-B-16": 41.09, "ViT-L-14-336": 395.22}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def usable_cores() -> int:
    """Host threads this process may really use: affinity mask and cgroup CPU quota, not the machine's core count."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


# ------------------------------------------------------------------------------------------- CPU reference arm
def run_cpu_reference(steps: int, warmup: int, batch: int = 32):
    """The reference's own CPU train step (config[0]: ViT-B-32 fp32, batch 32, all host threads), restated by the
    oracle port (oracle/clip_oracle.CpuTrainer: reference CLIPTask + train_step + AdamW + clamp)."""
    import torch
    from oracle import clip_oracle as O
    cores = usable_cores()
    torch.set_num_threads(cores)
    cfg = O.CONFIGS["ViT-B-32"]
    tr = O.CpuTrainer(cfg, seed=0)
    image, text = O.synthetic_batch(cfg, batch, seed=0)
    budget_s = float(os.environ.get("CLIPN_CPU_BASELINE_BUDGET_S", "40"))  # bounded sample: whole leg < ~1 min
    t_begin = time.perf_counter()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tr.step(image, text)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    timed = times[min(warmup, len(times) - 1):]
    dt = sum(timed) / len(timed)
    return {"value": batch / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"ViT-B-32 fp32 CPU train step, batch {batch}, {len(timed)} timed steps after "
                      f"{len(times) - len(timed)} warm-up, {cores} threads (oracle port of reference "
                      f"CLIPTask+train_step+AdamW)", "ms_per_step": dt * 1e3}


# ------------------------------------------------------------------------------------------- GPU arm
def logits_gemm_roofline(sig_time, sig_count, peak_tflops):
    """BASELINE.json's second figure: the fused gather + logits + log-sum-exp GEMMs of the loss forward.
    sig_time / sig_count map a GEMM signature (M, N, K, epilogue, a_mn, b_mn) to total ms / launches; the loss
    forward is the signature with the LSE epilogue (6). Returns None when no such launch was timed."""
    lse_sigs = [s for s in sig_time if s[3] == 6]
    n = sum(sig_count[s] for s in lse_sigs)
    ms = sum(sig_time[s] for s in lse_sigs)
    if n == 0 or ms <= 0:
        return None
    tf = sum(2.0 * s[0] * s[1] * s[2] * sig_count[s] for s in lse_sigs) / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "gemm_peer_kernel<BN,LSE> (BN = 256 for E <= 512, else 128; both directions in one "
                                        "launch: column tiles stationary in smem + online row LSE) + lse_combine",
            "achieved": tf, "peak": peak_tflops, "unit": "TFLOP/s", "frac": tf / peak_tflops, "launches_timed": n,
            "avg_launch_ms": ms / n, "shape_mnk": [list(s[:3]) for s in lse_sigs]}


def workload_config(model, batch, world, siglip=False, ckpt=False, optimizer="native", grad_sync="native"):
    """The `config` object of the native arm's JSON line."""
    if siglip:
        loss = "SigLipLoss, peer text/image blocks read in place (no neighbour exchange)"
    else:
        loss = "ClipLoss local_loss" + (" + gather_with_grad, gather fused into the logits GEMM" if world > 1 else " only")
    return {"workload": f"{model} bf16, local batch {batch}, {world}xB200, {loss}"
                        + (", grad-checkpointed blocks" if ckpt else ""),
            "global_batch": world * batch,
            "parallelism": f"dp{world}",
            "grad_sync": None if world == 1 else ("per-block NCCL all-reduce issued by the tower backward (libclipn arenas)"
                                                  if grad_sync == "native" else "torch DistributedDataParallel"),
            "optimizer": "AdamW multi-tensor, one launch (libclipn)" if optimizer == "native" else "AdamW fused (torch)",
            "cache": "inputs (>= 1.2 GB/step) and activations exceed the 126 MB L2; no explicit flush"}


def reference_config(batch, cores, gpus):
    """`config` of the reference arm: what the CPU leg really runs (the reference's CPU-runnable BASELINE config 0),
    named next to the native workload it is the baseline for."""
    return {"workload": f"ViT-B-32 fp32, batch {batch}, CPU {cores} threads (reference CLIPTask + train_step + AdamW, "
                        f"oracle port); baseline for the native arm at {gpus}xB200",
            "global_batch": batch, "parallelism": f"cpu{cores}", "optimizer": "AdamW (torch, CPU)"}


def oracle_rank0(all_i, all_t, scale, bias, siglip):
    """Rank 0's loss value and gradients (d image features, d text features, d logit_scale) by the oracle's restatement of
    the reference's multi-rank semantics, from every rank's features.  Runs in fp64 (falls back to fp32 if the device
    refuses): with the nearly parallel features of a fresh model d logit_scale is a difference of O(B) totals, and the fp32
    oracle's OWN error on it is 0.4-1 % at N = 32768 (against fp64; tests/test_host_logic.py pins the effect) — as large
    as the bar the kernels are held to."""
    import torch
    from oracle import clip_oracle as O

    def run(dt):
        ri, rt = [t.to(dt) for t in all_i], [t.to(dt) for t in all_t]
        if siglip:
            return tuple(O.siglip_loss_rank_grads(ri, rt, scale.to(dt), bias.to(dt), 0)[:4]) + (dt,)
        return tuple(O.clip_loss_rank_grads(ri, rt, scale.to(dt), 0, True, True)) + (dt,)
    try:
        return run(torch.float64)
    except RuntimeError:
        return run(torch.float32)


def parity_block(model, loss_fn, image, text, rank, world, siglip):
    """Oracle parity of the loss at the BENCHMARKED shape, outside the timed region: this step's features (bf16, from the
    native towers) go through the native loss (value + feature / logit_scale gradients) and, gathered with NCCL,
    through the oracle restatement of the reference's multi-rank semantics on rank 0 (fp32, on the GPU for speed —
    the oracle is device-agnostic torch code).  Tolerances: loss 1e-2 abs, feature gradients 1.5e-2 rel-L2,
    logit_scale gradient 2e-2 rel."""
    import torch
    import torch.distributed as dist
    from oracle import clip_oracle as O
    with torch.no_grad():
        out = model(image=image, text=text)
    # fp32 leaves holding the bf16 features' values: the loss value and the gradients come back in fp32 (a bf16 loss of
    # ~9 carries an OUTPUT rounding of up to 0.03, more than the tolerance) while the kernels see exactly the bf16 features
    fi = out["image_features"].detach().float().requires_grad_(True)
    ft = out["text_features"].detach().float().requires_grad_(True)
    sc = out["logit_scale"].detach().float().clone().requires_grad_(True)
    if siglip:
        lb = out["logit_bias"].detach().float().clone().requires_grad_(True)
        loss = loss_fn(fi, ft, sc, lb)
    else:
        loss = loss_fn(fi, ft, sc)
    loss.backward()
    if world > 1:
        all_i = [torch.empty_like(fi.detach()) for _ in range(world)]
        all_t = [torch.empty_like(ft.detach()) for _ in range(world)]
        dist.all_gather(all_i, fi.detach())
        dist.all_gather(all_t, ft.detach())
    else:
        all_i, all_t = [fi.detach()], [ft.detach()]
    res = None
    if rank == 0:
        ref, d_img, d_txt, d_scale, oracle_dt = oracle_rank0(all_i, all_t, sc.detach(), lb.detach() if siglip else None,
                                                             siglip)

        def rel(a, b):
            return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))
        res = {"loss": float(loss), "loss_ref": float(ref), "loss_abs_err": abs(float(loss) - float(ref)),
               "dfeat_rel_l2": max(rel(fi.grad, d_img), rel(ft.grad, d_txt)),
               "dscale_rel": abs(float(sc.grad) - float(d_scale)) / (abs(float(d_scale)) + 1e-30),
               "shape": {"world": world, "local_batch": int(fi.shape[0]), "embed": int(fi.shape[1])},
               "oracle": "oracle/clip_oracle.py %s (%s torch on cuda:0, rank 0)"
                         % ("siglip_loss_rank_grads" if siglip else "clip_loss_rank_grads local_loss+gather_with_grad",
                            "fp64" if oracle_dt == torch.float64 else "fp32"),
               "exchange": getattr(loss_fn, "exchange_mode", "local"),
               "tol": {"loss_abs": 1e-2 if not siglip else 2e-2 * abs(float(ref)) + 1e-3, "dfeat_rel_l2": 1.5e-2 if not siglip else 2e-2,
                       "dscale_rel": 2e-2}}
        res["ok"] = bool(res["loss_abs_err"] <= res["tol"]["loss_abs"] and res["dfeat_rel_l2"] <= res["tol"]["dfeat_rel_l2"]
                         and res["dscale_rel"] <= res["tol"]["dscale_rel"])
    model.zero_grad(set_to_none=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="local batch per GPU (BASELINE config: 4096)")
    ap.add_argument("--model", default="ViT-B-32")
    ap.add_argument("--grad-checkpointing", action="store_true", help="BASELINE config 4 (ViT-L-14-336) runs with it")
    ap.add_argument("--siglip", action="store_true",
                    help="BASELINE config 5: SigLipLoss, init_logit_scale=ln 10, init_logit_bias=-10 (main.py:259-261)")
    ap.add_argument("--optimizer", default="native", choices=["native", "torch"],
                    help="native: libclipn multi-tensor AdamW (one launch); torch: torch.optim.AdamW(fused=True)")
    ap.add_argument("--grad-sync", default="native", choices=["native", "ddp"],
                    help="N > 1: native = NativeCLIP.enable_grad_sync() (per-block all-reduce issued by the tower backward); "
                         "ddp = torch DistributedDataParallel as the reference wraps it (base_task.py:227)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity block (outside the timed region)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        warm = max(1, min(args.warmup, 2))
        steps = max(1, min(args.steps, 5))
        r = run_cpu_reference(steps, warm)
        print(json.dumps({
            "impl": "reference", "metric": "image-text pairs/sec (full train step)", "value": r["value"],
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # what really ran: the reference's CPU-runnable config 0 (ViT-B-32 fp32, batch 32) on the host threads
            "config": reference_config(32, r["cores"], max(1, args.gpus)),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist
    from open_clip_b200 import _lib, ops
    from open_clip_b200.loss import NativeClipLoss, NativeSigLipLoss
    from open_clip_b200.model import create_model
    from oracle import clip_oracle as O  # checker only: cpu_baseline leg, parity block, synthetic data recipe

    _lib.lib()  # fail loudly if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = run_cpu_reference(steps=3, warmup=2)

    B = args.batch
    torch.manual_seed(0)
    mkw = dict(init_logit_scale=math.log(10), init_logit_bias=-10.0) if args.siglip else {}
    model = create_model(args.model, output_dict=True, device=dev, **mkw)
    if args.grad_checkpointing:
        model.set_grad_checkpointing(True)
    step_gflop = 3 * FWD_GFLOP_BY_MODEL.get(args.model, FWD_GFLOP_PER_PAIR)
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    if args.siglip:
        loss_fn = NativeSigLipLoss(rank=rank, world_size=world)
    else:
        loss_fn = NativeClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
    named = list(model.named_parameters())
    no_wd = model.no_weight_decay()
    decay = [p for n, p in named if p.ndim > 1 and n not in no_wd]
    no_decay = [p for n, p in named if not (p.ndim > 1 and n not in no_wd)]
    groups = [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.2}]
    if args.optimizer == "native":
        from open_clip_b200.optim import NativeAdamW
        opt = NativeAdamW(groups, lr=5e-4, betas=(0.9, 0.98), eps=1e-6)
    else:
        opt = torch.optim.AdamW(groups, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, fused=True)
    module = model
    if world > 1:
        if args.grad_sync == "native":
            model.enable_grad_sync()
        else:
            module = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], bucket_cap_mb=200)

    cfg = O.CONFIGS[args.model]
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    pool = 3
    d_images = [torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=gen, device=dev, dtype=torch.bfloat16)
                for _ in range(pool)]
    d_texts = []
    for _ in range(pool):
        t = torch.randint(1, cfg.t_vocab - 1, (B, cfg.t_ctx), generator=gen, device=dev)
        t[:, -1] = cfg.t_vocab - 1
        d_texts.append(t)

    def train_step(image, text):
        opt.zero_grad(set_to_none=True)
        out = module(image=image, text=text)
        if args.siglip:
            loss = loss_fn(out["image_features"], out["text_features"], out["logit_scale"], out["logit_bias"])
        else:
            loss = loss_fn(out["image_features"], out["text_features"], out["logit_scale"])
        loss.backward()
        opt.step()
        with torch.no_grad():
            model.logit_scale.clamp_(0, math.log(100))
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_steps, nsteps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(nsteps)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / nsteps

    # ---------------- device-resident run ("value") ----------------
    def resident(n):
        for i in range(n):
            train_step(d_images[i % pool], d_texts[i % pool])

    resident(args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # dominant kernel = the tcgen05 GEMM family (2/3 of the step): CUDA events around EVERY GEMM launch of the timed
    # region, on the launching stream; reported per signature (largest total first) and as a family aggregate
    ops.PROFILE_KEY = "all"
    ops.PROFILE_EVENTS.clear()
    ops.LAUNCHES = 0
    ops.stage_timing(True)   # library-side events between the gather / GEMM / combine kernels of the fused loss forward
    ms_step = timed(resident, args.steps)
    stage = ops.stage_times()
    ops.stage_timing(False)
    launches = ops.LAUNCHES // max(args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    sig_time, sig_count = {}, {}
    for a, b_, sig in ops.PROFILE_EVENTS:
        sig_time[sig] = sig_time.get(sig, 0.0) + a.elapsed_time(b_)
        sig_count[sig] = sig_count.get(sig, 0) + 1
    ops.PROFILE_KEY = None
    peaks = load_peaks()
    pairs_per_s = world * B / (ms_step * 1e-3)

    # ---------------- end-to-end run (pinned host buffers, H2D inside the timed region, D2H of the loss) -------------
    e2e = None
    if not args.no_e2e:
        h_img = [d_images[i].cpu().pin_memory() for i in range(2)]
        h_txt = [d_texts[i].cpu().pin_memory() for i in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        stage_img = [torch.empty_like(d_images[0]) for _ in range(2)]
        stage_txt = [torch.empty_like(d_texts[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        host_loss = torch.empty(1, dtype=torch.float32).pin_memory()

        def prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])
                stage_img[s].copy_(h_img[i % 2], non_blocking=True)
                stage_txt[s].copy_(h_txt[i % 2], non_blocking=True)
                ready[s].record(copy_stream)

        def e2e_steps(n):
            for s in range(2):
                consumed[s].record(torch.cuda.current_stream())
            prefetch(0)
            for i in range(n):
                s = i % 2
                if i + 1 < n:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[s])
                loss = train_step(stage_img[s], stage_txt[s])
                consumed[s].record(torch.cuda.current_stream())
                host_loss.copy_(loss.detach().float().reshape(1), non_blocking=True)  # D2H of the step's result
            torch.cuda.current_stream().synchronize()

        e2e_steps(2)
        ms_e2e = timed(e2e_steps, args.steps)
        h2d = h_img[0].numel() * h_img[0].element_size() + h_txt[0].numel() * h_txt[0].element_size()
        e2e = {"value": world * B / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}

    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(model, loss_fn, d_images[0], d_texts[0], rank, world, args.siglip)
        except Exception as exc:  # the checker must never cost the measurement: report, do not die
            parity = {"ok": False, "error": "%s: %s" % (type(exc).__name__, exc)} if rank == 0 else None

    if rank == 0:
        epi_names = ["STORE", "BIAS_GELU", "BIAS_RESID", "DGELU", "ACCUM_F32(split-K wgrad)", "STORE_F32", "LSE",
                     "CLIP_DLOGITS", "SIGLIP", "BIAS_GELU_GRAD", "MUL_AUX"]
        fam_flops = sum(2.0 * s[0] * s[1] * s[2] * sig_count[s] for s in sig_time)
        fam_ms = sum(sig_time.values())
        top = max(sig_time, key=sig_time.get) if sig_time else None
        top_ms = sig_time[top] / sig_count[top] if top else float("nan")
        flops_launch = 2.0 * top[0] * top[1] * top[2] if top else 0.0
        avg_ms = top_ms
        achieved = flops_launch / (avg_ms * 1e-3) / 1e12 if top else None
        gemm_ms = [0] * (sig_count[top] if top else 0)
        top_name = ("gemm_tc2_kernel<256,%s> M=%d N=%d K=%d%s" % (epi_names[top[3]], top[0], top[1], top[2],
                    " (MN-major operands)" if top[4] else "")) if top else "n/a"
        # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of THIS kernel signature, taken from an
        # `ncu --set full` capture of the current binary and recorded in profiles/ncu_traffic.json by
        # tools/ncu_traffic.py ({"<epi>,<N>,<K>": {"m": M, "bytes": B, "source": "..."}}; scaled linearly in M).
        # null when no capture of the timed signature exists — never a number from a different kernel.
        traffic, traffic_src = None, None
        try:
            tbl = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tbl.get("%d,%d,%d" % (top[3], top[1], top[2])) if top else None
            if ent:
                traffic, traffic_src = ent["bytes"] * top[0] / ent["m"], ent.get("source")
        except (OSError, ValueError, KeyError):
            pass
        logits_roofline = logits_gemm_roofline(sig_time, sig_count, peaks["bf16_sustained"])
        if logits_roofline is not None and stage[0] > 0 and not args.siglip:
            # the launch split into its kernels: peer gather (NVLink reads, W > 1) | tcgen05 GEMM + online LSE | combine
            flops = 2.0 * 2 * B * world * B * model.embed_dim
            logits_roofline["stages_ms"] = {"calls": stage[0], "peer_gather": stage[1], "gemm": stage[2], "combine": stage[3]}
            logits_roofline["gemm_kernel_frac"] = flops / (stage[2] * 1e-3) / 1e12 / peaks["bf16_sustained"] if stage[2] > 0 else None
        out = {
            "metric": "image-text pairs/sec (full train step)", "value": pairs_per_s, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.model, B, world, args.siglip, args.grad_checkpointing, args.optimizer,
                                      args.grad_sync),
            "roofline": {"bound": "tensor", "kernel": top_name,
                         "share_of_step": (sig_time[top] / args.steps) / ms_step if top else None,
                         "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["bf16_sustained"]) if achieved else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         # operands once + the epilogue's bytes per output element (bf16 tensors 2 B: BIAS_GELU(_GRAD) write
                         # two, MUL_AUX / BIAS_RESID read one and write one, DGELU reads one and writes two; fp32 outputs 4 B)
                         "algorithmic_bytes": (2.0 * (top[0] * top[2] + top[1] * top[2]) + top[0] * top[1] *
                                               {1: 4, 2: 4, 3: 6, 4: 4, 5: 4, 9: 4, 10: 4}.get(top[3], 2)) if top else None,
                         "launches_timed": len(gemm_ms), "avg_launch_ms": avg_ms,
                         "flops_per_launch": flops_launch, "peak_source": peaks["source"] + ", sustained"},
            "roofline_gemm_family": {"bound": "tensor", "achieved": fam_flops / (fam_ms * 1e-3) / 1e12 if fam_ms else None,
                                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                                     "frac": fam_flops / (fam_ms * 1e-3) / 1e12 / peaks["bf16_sustained"] if fam_ms else None,
                                     "launches_timed": len(ops.PROFILE_EVENTS),
                                     "share_of_step": (fam_ms / args.steps) / ms_step},
            "roofline_step": {"bound": "tensor", "achieved": pairs_per_s / world * step_gflop / 1e3,
                              "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                              "frac": pairs_per_s / world * step_gflop / 1e3 / peaks["bf16_sustained"],
                              "flops_per_pair": step_gflop * 1e9},
            "roofline_logits_gemm": logits_roofline,
            # every GEMM signature of the timed region, largest total first (same live CUDA-event timings)
            "roofline_gemm_signatures": [
                {"epilogue": epi_names[s[3]], "m": s[0], "n": s[1], "k": s[2], "mn_major": bool(s[4]),
                 "launches_per_step": sig_count[s] / args.steps, "avg_launch_ms": sig_time[s] / sig_count[s],
                 "share_of_step": (sig_time[s] / args.steps) / ms_step,
                 "frac": 2.0 * s[0] * s[1] * s[2] / (sig_time[s] / sig_count[s] * 1e-3) / 1e12 / peaks["bf16_sustained"]}
                for s in sorted(sig_time, key=sig_time.get, reverse=True)[:12]],
            "gpu_launches": launches,
            "clocks": clocks,
            "hbm_peak_allocated_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
        }
        if e2e is not None:
            out["e2e"] = e2e
        if cpu_base is not None:
            out["cpu_baseline"] = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if parity is not None:
            out["parity"] = parity
        print(json.dumps(out), flush=True)
    parity_failed = parity is not None and not parity["ok"]
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity_failed:
        # the JSON line above already carries "parity": {"ok": false}; the timing itself is valid, so the exit status
        # stays 0 unless CLIPN_BENCH_STRICT_PARITY=1 asks for a hard failure
        sys.stderr.write("bench.py: PARITY FAILED against the oracle: %s\n" % json.dumps(parity))
        if os.environ.get("CLIPN_BENCH_STRICT_PARITY") == "1":
            sys.exit(3)


if __name__ == "__main__":
    main()
