#!/usr/bin/env python
"""Benchmark of the fcd_b200 hot path.   python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): MS_DSA_NET (feature_size 16, project_size 64, 'parallel' DSA, reference dropout
p=0.1 active), bf16 compute, synthetic 2-channel 128^3 patches, batch 2 per GPU, DiceCELoss, AdamW(1e-4, wd 1e-5);
one "step" = forward + fused loss + backward + gradient all-reduce (N>1) + optimizer step.  Weak scaling.

  value : patches/s with the inputs resident in HBM (whole job, all ranks), CUDA-event timed, max over ranks.
  e2e   : the same metric through the public API with HOST (pinned) inputs: the H2D copy of the batch and the D2H
          read of the loss are inside every timed step.
  roofline     : the tcgen05 conv kernel (and the whole conv family) measured with CUDA events inside one eager step.
  cpu_baseline : the CPU oracle (oracle/, kind "port": the reference's MONAI dependency is not installable here)
                 timed on the host cores on a bounded sample (rank 0, N=1 only).
  aux          : sliding-window inference vols/s on a synthetic 256x256x192 volume (configs[4]).

`--impl reference` times the reference's CPU implementation of the path (the oracle port) on the host cores.
L2 note: every step streams > 3 GB of activations, far above the 126 MB L2 (config.l2 = "inputs_exceed_l2").
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms_dsa_net_train_patches_per_s"
UNIT = "patches/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2, help="patches per GPU")
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--model", default="ms_dsa_net")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one gradient all-reduce after backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=2, help="patches timed for the CPU baseline")
    ap.add_argument("--loss", default=None, help="DiceCELoss (default) | DiceFocalLoss | DiceLoss")
    ap.add_argument("--tv", type=float, default=0.0, help="tv_loss_weight (configs[3] uses 0.1)")
    ap.add_argument("--torch-adamw", action="store_true", help="step with torch.optim.AdamW(fused=True) instead of FusedAdamW")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip timing the stock PyTorch (cuDNN/cuBLAS) bf16/fp16-autocast step on the same GPU")
    return ap.parse_args()


def loss_name(args):
    if args.loss:
        return args.loss
    return "DiceFocalLoss" if args.model.startswith("segresnet") and "dsa" not in args.model else "DiceCELoss"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _kernel_errors():
    """Status word of the bounded mbarrier waits in the tcgen05 kernels (0 = no wait ever timed out) + its record."""
    from fcd_b200 import _lib
    st = _lib.status()
    return {"word": st["word"]} if st["word"] == 0 else st


def gpu_reference(model_type, patch, batch, loss_over, dev, steps=4, warm=2):
    """The existing Blackwell implementation of the path (SURVEY 2.1 / 8d "the kernel to beat"): the SAME network
    arithmetic through stock PyTorch -- cuDNN / cuBLAS eager kernels -- on this GPU, under bf16 autocast and under fp16
    autocast (+ GradScaler) as the reference trains (train.py:328,373): forward + loss + backward + fused AdamW, batch
    `batch`.  The network is the functional restatement of the reference modules (oracle/nets.py; the reference's own
    files need MONAI, which cannot be installed here) and runs WITHOUT dropout, which favours it slightly."""
    from oracle import losses as olosses
    from oracle import nets as onets
    from oracle import synth
    import fcd_b200
    params = fcd_b200.get_default_params()
    params.update(model_type=model_type, patch_size=(patch,) * 3)
    params.update(loss_over)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    spec = synth.spec_of(model.state_dict())
    del model
    sd = synth.synthetic_state_dict(spec, seed=1)
    fk = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    x = synth.image(batch, 2, patch, seed=3).to(dev)
    y = synth.label(batch, patch, seed=5).to(dev)
    noise = torch.randn((batch, 256), device=dev)
    out = {}
    for name, dtype in (("bf16_autocast", torch.bfloat16), ("fp16_autocast", torch.float16)):
        try:
            leaf = {k: (v.to(dev).clone().requires_grad_(True) if k in fk else v.to(dev).clone()) for k, v in sd.items()}
            opt = torch.optim.AdamW([leaf[k] for k in fk], lr=1e-4, weight_decay=1e-5, fused=True)
            scaler = torch.amp.GradScaler("cuda", enabled=dtype == torch.float16)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(warm + steps):
                if i == warm:
                    torch.cuda.synchronize()
                    e0.record()
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=dtype):
                    o = onets.forward(model_type, leaf, x, True, {}, noise)
                    vae = None
                    if isinstance(o, tuple):
                        o, vae = o
                    loss = olosses.combined_loss(params, o, y)
                    if vae is not None:
                        loss = loss + params["loss_vae_weight"] * vae
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
            del leaf, opt
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    out["what"] = (f"oracle/nets.py functional {model_type} through stock torch {torch.__version__} eager (cuDNN "
                   f"{torch.backends.cudnn.version()}), fwd + loss + bwd + fused AdamW, batch {batch}, {steps} steps after "
                   f"{warm} warm-up, no dropout, no CUDA graph")
    return out


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference(model_type, patch, n_patches, warm=1, loss_over=None):
    """The reference's CPU implementation of the path: the oracle port (reference network files cannot travel, and
    their MONAI dependency is not installable here) -- fp32, all host threads, batch 1, fwd + DiceCE + bwd."""
    from oracle import losses as olosses
    from oracle import nets as onets
    from oracle import synth
    import fcd_b200
    torch.set_num_threads(os.cpu_count() or 1)
    params = fcd_b200.get_default_params()
    params.update(model_type=model_type, patch_size=(patch,) * 3, loss="DiceCELoss")
    params.update(loss_over or {})
    # parameter names / shapes come from the product's module tree (identical to the reference's); values synthetic
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    spec = synth.spec_of(model.state_dict())
    del model
    sd = synth.synthetic_state_dict(spec, seed=1)
    fk = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    leaf = {k: (v.requires_grad_(True) if k in fk else v) for k, v in sd.items()}
    x = synth.image(1, 2, patch, seed=3)
    y = synth.label(1, patch, seed=5)
    times = []
    for i in range(warm + n_patches):
        t0 = time.perf_counter()
        out = onets.forward(model_type, leaf, x, True, {})
        loss = olosses.combined_loss(params, out, y)
        loss.backward()
        for k in fk:
            leaf[k].grad = None
        times.append(time.perf_counter() - t0)
    t = times[warm:]
    return len(t) / sum(t), sum(t) / len(t)


def cpu_sliding_window(model_type, patch, cores):
    """CPU baseline of configs[4] on a bounded sample: ONE predictor call of 2 windows (the reference's sw_batch_size,
    train.py:159) of the 18 through the oracle network in eval mode, scaled to 9 calls, plus the reference's scipy
    post-processing (utils_common.py:10-33) measured on a 128^3 crop and scaled by volume."""
    from oracle import inferer as oinf
    from oracle import nets as onets
    from oracle import synth
    import fcd_b200
    torch.set_num_threads(cores)
    params = fcd_b200.get_default_params()
    params.update(model_type=model_type, patch_size=(patch,) * 3)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=1)
    del model
    x = synth.image(2, 2, patch, seed=3)
    with torch.no_grad():
        onets.forward(model_type, sd, x[:1], False, {})
        t0 = time.perf_counter()
        out = onets.forward(model_type, sd, x, False, {})
        t_call = time.perf_counter() - t0
    mask = (synth.label(1, 128, seed=31, n_blobs=4)[0, 0] > 0).numpy().astype("float32")
    t0 = time.perf_counter()
    oinf.post_process_segment(mask, 50)
    t_pp = (time.perf_counter() - t0) * (256 * 256 * 192) / 128 ** 3
    sec = 9 * t_call
    return {"value": 1.0 / sec, "unit": "vols/s", "cores": cores, "kind": "port",
            "sample": f"1 of 9 predictor calls (2 windows of {patch}^3, eval, fp32, {cores} threads: {t_call:.1f} s) scaled to "
                      f"the 18 windows; blend/finalize excluded",
            "post_process_s_per_vol": t_pp, "post_process_sample": "scipy post_process_segment on a 128^3 crop, scaled by volume"}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    k = max(1, min(args.steps, 3))
    pps, sec = cpu_reference(args.model, args.patch, k, warm=1 if args.warmup > 0 else 0,
                             loss_over=dict(loss=loss_name(args), tv_loss_weight=args.tv))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} train step (fwd + {loss_name(args)} + bwd), 2ch {args.patch}^3, batch 1, CPU"},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{k} patches of batch 1 after 1 warm-up, torch {torch.__version__} CPU fp32"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import fcd_b200
    from fcd_b200 import _lib, parallel, synthetic
    from fcd_b200.inferers import sliding_window_inference

    rank, local, world = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pk = peaks()

    params = fcd_b200.get_default_params()
    loss_over = dict(loss=loss_name(args), tv_loss_weight=args.tv)
    params.update(model_type=args.model, patch_size=(args.patch,) * 3, **loss_over)
    torch.manual_seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    model.apply(synthetic.initialize_weights)
    model = model.to(dev).train()
    loss_fn = fcd_b200.CombinedLoss(params, dev)
    if args.torch_adamw:
        opt = torch.optim.AdamW(model.parameters(), lr=params["lr"], weight_decay=params["weight_decay"], fused=True)
    else:       # one fcd_adamw_multi launch for all parameter tensors (torch's fused AdamW: ~10 launches)
        from fcd_b200.optim import FusedAdamW
        opt = FusedAdamW(model.parameters(), lr=params["lr"], weight_decay=params["weight_decay"])
    reducer = parallel.GradAllReducer(model.parameters(), overlap=not args.no_overlap)
    reducer.sync_params()
    B = args.batch
    n_pool = 4
    host = [synthetic.make_batch(B, 2, args.patch, seed=100 * rank + i, pin=True) for i in range(n_pool)]
    sx = torch.empty_like(host[0][0], device=dev)
    sy = torch.empty_like(host[0][1], device=dev)
    sx.copy_(host[0][0])
    sy.copy_(host[0][1])
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    static = {}

    def fwd_bwd():
        out = model(sx)
        vae = None
        if isinstance(out, (tuple, list)):
            out, vae = out
        loss = loss_fn(out, sy)
        if vae is not None:
            loss = loss + params["loss_vae_weight"] * vae
        loss.backward()
        return loss.detach()

    # ---- eager warm-up (also configures every kernel's shared-memory attribute) + launch count + roofline pass
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        fwd_bwd()
        reducer.allreduce()
        opt.step()
    torch.cuda.synchronize()
    # per-launch durations for the roofline object: one eager step with CUDA events around every C-ABI call, with the
    # stream overlap (side-stream weight gradients, branch streams) switched OFF so that every launch is timed alone
    # on its launching stream -- inside the overlapped step a launch shares the SMs with other streams' kernels
    from fcd_b200 import ops as _ops
    saved_overlap = (_ops.WGRAD_OVERLAP, _ops.BRANCH_OVERLAP)
    _ops.WGRAD_OVERLAP = _ops.BRANCH_OVERLAP = False
    reducer.enabled = False
    prof = _lib.Profiler()
    _lib.set_profiler(prof)
    opt.zero_grad(set_to_none=True)
    fwd_bwd()
    _lib.set_profiler(None)
    agg = prof.summary()
    gpu_launches = prof.launches + (0 if args.torch_adamw else 1)      # + the one-launch optimizer step
    _ops.WGRAD_OVERLAP, _ops.BRANCH_OVERLAP = saved_overlap
    reducer.enabled = True
    reducer.allreduce()
    opt.step()

    # ---- CUDA graph of forward + loss + backward (static input buffers, static gradient tensors)
    graph = None
    if not args.no_graph:
        try:
            opt.zero_grad(set_to_none=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    opt.zero_grad(set_to_none=True)
                    fwd_bwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            opt.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=_ops.compute_stream(dev)):
                static["loss"] = fwd_bwd()
            torch.cuda.synchronize()
        except Exception as e:  # keep the bench alive; the JSON line records that the step ran eagerly
            graph = None
            static["graph_error"] = f"{type(e).__name__}: {e}"[:200]
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
            loss = static["loss"]
        else:
            opt.zero_grad(set_to_none=True)
            loss = fwd_bwd()
        reducer.allreduce(early_in_graph=graph is not None and reducer.early_captured)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # e2e input pipeline: the pinned-host batch of step i+1 is copied to a staging buffer on a copy stream while step
    # i computes (what a DataLoader with pin_memory + non_blocking copies does); every step still pays its own H2D
    # copy (inside the timed region), a device-side staging->input copy, and the D2H read of its loss.
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [(torch.empty_like(sx), torch.empty_like(sy)) for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(it):
        k = it % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])                            # staging slot k no longer being read
            stage[k][0].copy_(host[it % n_pool][0], non_blocking=True)     # H2D of batch `it` (pinned)
            stage[k][1].copy_(host[it % n_pool][1], non_blocking=True)
            staged[k].record(copy_stream)

    def e2e_step(it):
        k = it % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(staged[k])
        sx.copy_(stage[k][0], non_blocking=True)
        sy.copy_(stage[k][1], non_blocking=True)
        consumed[k].record(cur)
        prefetch(it + 1)                                                   # overlaps this step's compute
        loss_host.copy_(step(), non_blocking=True)                         # D2H of the step's loss ...
        cur.synchronize()                                                  # ... read by the host every step (train.py:382)
        return float(loss_host)

    def timed(n_warm, n_steps, e2e):
        it = 0
        if e2e:
            for k in range(2):
                consumed[k].record(torch.cuda.current_stream())
            prefetch(0)
        for _ in range(n_warm):
            if e2e:
                e2e_step(it)
            else:
                step()
            it += 1
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            if e2e:
                e2e_step(it)
            else:
                step()
            it += 1
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms

    W = max(args.warmup, 3)
    graph_ok, graph_err = graph is not None, static.get("graph_error")
    with ClockSampler(local) as clk:
        ms_dev = timed(W, args.steps, e2e=False)
        ms_e2e = timed(W, args.steps, e2e=True)
    clocks = clk.summary()
    patches = B * world * args.steps
    value = patches / (ms_dev / 1e3)
    e2e_value = patches / (ms_e2e / 1e3)
    final_loss = float(loss_host)
    errs = _kernel_errors()
    if errs["word"]:
        raise RuntimeError(f"tcgen05 pipeline time-out during the training steps: {errs}")

    # ---- roofline from the profiled eager step (CUDA events around every C-ABI call on the launching stream, stream
    # overlap off so that every launch is timed alone).  Per-family table; the DOMINANT family is the one with the largest
    # share of the serialized step, and the `roofline` object describes it (bound, achieved, peak, frac).
    FAMILIES = [
        ("tcgen05 conv3x3x3 fwd/dgrad", "tensor", ("fcd_conv3_tcf", "fcd_conv3_tc")),
        ("tcgen05 deep-level GEMM conv (+split-K reduce)", "tensor", ("fcd_conv_gemm_tc", "fcd_splitk_reduce")),
        ("tcgen05 weight gradients (+reduce)", "tensor", ("fcd_wgrad3_tc", "fcd_wgrad_gemm_tc", "fcd_wgrad_reduce")),
        ("TMA + tcgen05 row GEMM (1x1x1 conv / linear / deconv k2s2)", "hbm", ("fcd_rowgemm",)),
        ("mma.sync conv / linear / deconv / pointwise", "tensor", ("fcd_igemm", "fcd_igemm_splitk", "fcd_pw_conv",
                                                                   "fcd_wgrad", "fcd_pack_weight", "fcd_pack_weight_batched")),
        ("norm + activation + residual (IN/BN/GN)", "hbm", ("fcd_norm_apply", "fcd_norm_stats", "fcd_norm_finalize",
                                                            "fcd_norm_bwd")),
        ("DSA attention + LayerNorm", "hbm", ("fcd_dsa_fwd", "fcd_dsa_bwd", "fcd_dsa_bwd_ef", "fcd_ln_fwd", "fcd_ln_bwd")),
        ("loss + output conv", "hbm", ("fcd_loss_fwd", "fcd_loss_bwd", "fcd_outconv_fwd", "fcd_outconv_bwd",
                                       "fcd_mse_fwd", "fcd_mse_bwd")),
    ]
    step_ms_eager = sum(v["ms"] for v in agg.values())
    tf = (lambda fl, ms: fl / (ms * 1e-3) / 1e12 if ms > 0 else None)
    gbs = (lambda nb, ms: nb / (ms * 1e-3) / 1e9 if ms > 0 else None)
    seen = set()
    table = []
    for label, bound, names in FAMILIES:
        sel = {k: v for k, v in agg.items() if k.split(":")[0] in names}
        seen.update(sel)
        ms_, fl_, nb_, calls_ = (sum(v[f] for v in sel.values()) for f in ("ms", "flops", "bytes", "calls"))
        row = {"family": label, "bound": bound, "ms": round(ms_, 3), "share_of_step": ms_ / step_ms_eager if step_ms_eager else None,
               "calls": calls_}
        if bound == "tensor" and fl_ > 0 and ms_ > 0:
            row.update(achieved=tf(fl_, ms_), unit="TFLOP/s", frac=tf(fl_, ms_) / pk["tf_sust"])
        elif bound == "hbm" and nb_ > 0 and ms_ > 0:
            row.update(achieved=gbs(nb_, ms_), unit="GB/s", frac=gbs(nb_, ms_) / pk["hbm"])
        table.append(row)
    rest = {k: v for k, v in agg.items() if k not in seen}
    rest_ms = sum(v["ms"] for v in rest.values())
    table.append({"family": "other (pool, upsample, copies, adds, dropout masks, layout)", "bound": "hbm", "ms": round(rest_ms, 3),
                  "share_of_step": rest_ms / step_ms_eager if step_ms_eager else None,
                  "calls": sum(v["calls"] for v in rest.values())})
    dom = max(table[:-1], key=lambda r: r["ms"])
    dom_names = next(n for l, _, n in FAMILIES if l == dom["family"])
    # the dominant family's largest single launch (by algorithmic work) is the kernel the roofline object is quoted on
    big = None
    for name, tag, ev0, ev1, fl, nb in prof.records:
        w = fl if dom["bound"] == "tensor" else nb
        if name in dom_names and w > 0 and (big is None or w > big[1]):
            big = (name + (":" + tag if tag else ""), w, ev0.elapsed_time(ev1))
    # measured DRAM traffic of that kind of launch, from the committed ncu capture (per launch), if there is one
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and big is not None:
        with open(tpath) as f:
            tj = json.load(f)
        hit = tj.get(big[0].split(":")[0])
        if hit:
            traffic, traffic_src = hit.get("dram_bytes"), hit
    roof = {"bound": dom["bound"], "kernel": dom["family"], "achieved": dom.get("achieved"),
            "peak": pk["tf_sust"] if dom["bound"] == "tensor" else pk["hbm"], "unit": dom.get("unit"),
            "frac": dom.get("frac"), "traffic": traffic, "traffic_source": traffic_src,
            "share_of_step": dom["share_of_step"], "launches": dom["calls"],
            "largest_launch": None if big is None else {
                "call": big[0], ("algorithmic_flops" if dom["bound"] == "tensor" else "algorithmic_bytes"): big[1],
                "ms": round(big[2], 4),
                "achieved": (tf(big[1], big[2]) if dom["bound"] == "tensor" else gbs(big[1], big[2])),
                "frac": (tf(big[1], big[2]) / pk["tf_sust"] if dom["bound"] == "tensor" else gbs(big[1], big[2]) / pk["hbm"])},
            "peak_source": pk["source"] + (" (sustained bf16)" if dom["bound"] == "tensor" else " (copy bandwidth)"),
            "timing": "CUDA events around each launch in one serialized eager step (stream overlap off); the kernels' "
                      "share of the step agrees with the ncu launch list under profiles/",
            "serialized_step_ms": round(step_ms_eager, 3),
            "families": table,
            "top_calls_ms": {k: round(v["ms"], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:12]}}
    whole_flops = sum(v["flops"] for v in agg.values())
    roof["whole_step"] = {"algorithmic_tflop": whole_flops / 1e12, "achieved": tf(whole_flops, ms_dev / args.steps),
                          "frac_of_tensor_peak": tf(whole_flops, ms_dev / args.steps) / pk["tf_sust"]}

    # ---- sliding-window whole-volume inference (configs[4]): 256x256x192, roi 128, overlap 0.5 -> 18 windows.
    # value : volume resident in HBM -> uint8 label map in HBM (all ranks hold it);  e2e : pinned HOST volume -> HOST label
    # map (rank 0 copies the volume once and broadcasts it over NVLink at N > 1; D2H of the label map on every rank).
    # N > 1: windows dealt round-robin, fp32 partial volumes reduce-scattered along D, slab finalize, label all-gather.
    aux = {}
    if not args.no_infer and args.model in ("ms_dsa_net", "ms_dsa_net_ps", "baseunet", "segresnet"):
        try:
            from fcd_b200.inferers import post_process_segment
            del graph
            static.clear()
            opt.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            model.eval()
            vshape = (1, 2, 256, 256, 192)
            vol = torch.randn(vshape, generator=torch.Generator().manual_seed(7)).pin_memory()
            vol_d = torch.empty_like(vol, device=dev)
            vol_d.copy_(vol)
            lab_h = torch.empty((1, 1) + vshape[2:], dtype=torch.uint8).pin_memory()
            # sw_batch_size is free in configs[4]: every rank pushes all of its windows through ONE
            # forward (18 at N=1, 3 at N=8) -- the deep levels are latency-bound, so they amortise over the batch
            sw_bs = 18

            def infer_dev():
                _, lab = sliding_window_inference(vol_d, args.patch, sw_bs, model, overlap=0.5, label_mode="argmax",
                                                  shard=world > 1, return_logits=False)
                return lab

            def infer_e2e_serial():
                if rank == 0:
                    vol_d.copy_(vol, non_blocking=True)                        # H2D of the volume (pinned), once per job
                if world > 1:
                    torch.distributed.broadcast(vol_d, 0)                      # NVLink, not 8 PCIe copies
                lab = infer_dev()
                lab_h.copy_(lab, non_blocking=True)                            # D2H of the label map (pinned)
                torch.cuda.current_stream().synchronize()
                return lab

            # N = 1: a two-deep pipeline, as a reader thread + pinned buffers give any inference server: the H2D of
            # volume k+1 and the D2H of label map k-1 run on a copy stream while volume k computes.  Every volume still
            # pays its own 100 MB H2D and 12.6 MB D2H inside the timed region; they overlap the tensor work instead of
            # preceding / following it.
            copy_stream = torch.cuda.Stream(device=dev)
            pipe = {"k": 0, "h2d": [None, None], "free": [None, None], "d2h": [None, None], "lab": [None, None]}
            stage = [vol_d, torch.empty_like(vol_d)] if world == 1 else None
            lab_hs = [lab_h, torch.empty_like(lab_h).pin_memory()] if world == 1 else None

            def _prefetch(k):
                b = k & 1
                with torch.cuda.stream(copy_stream):
                    if pipe["free"][b] is not None:
                        copy_stream.wait_event(pipe["free"][b])              # the forward that read this buffer is done
                    stage[b].copy_(vol, non_blocking=True)                     # H2D of volume k (pinned)
                    pipe["h2d"][b] = torch.cuda.Event()
                    pipe["h2d"][b].record(copy_stream)

            def infer_e2e_pipelined():
                k = pipe["k"]
                b = k & 1
                cur = torch.cuda.current_stream()
                if pipe["h2d"][b] is None:
                    _prefetch(k)                                               # first call: nothing was prefetched yet
                cur.wait_event(pipe["h2d"][b])
                pipe["h2d"][b] = None
                _prefetch(k + 1)
                _, lab = sliding_window_inference(stage[b], args.patch, sw_bs, model, overlap=0.5, label_mode="argmax",
                                                  return_logits=False)
                pipe["free"][b] = torch.cuda.Event()
                pipe["free"][b].record(cur)
                if pipe["d2h"][b] is not None:
                    pipe["d2h"][b].synchronize()                               # label map k-2 has long arrived: reuse its buffer
                pipe["lab"][b] = lab                                           # keeps the device label map alive for the copy
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(pipe["free"][b])
                    lab_hs[b].copy_(lab, non_blocking=True)                    # D2H of label map k (pinned)
                    pipe["d2h"][b] = torch.cuda.Event()
                    pipe["d2h"][b].record(copy_stream)
                if pipe["d2h"][b ^ 1] is not None:
                    pipe["d2h"][b ^ 1].synchronize()                           # the host consumes label map k-1 here
                pipe["k"] = k + 1
                return lab

            def e2e_drain():
                """End of the timed region: the last label map must have reached the host."""
                for ev in pipe["d2h"]:
                    if ev is not None:
                        torch.cuda.current_stream().wait_event(ev)
                        ev.synchronize()

            infer_e2e = infer_e2e_pipelined if world == 1 else infer_e2e_serial

            def timed_vols(fn, n_vol, drain=None):
                import gc
                for _ in range(3):
                    fn()
                barrier()
                gc.collect()
                gc.disable()      # a generational collection inside the loop showed up as a 7-80 ms stall on one volume
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_vol + 1)]
                evs[0].record()
                for k in range(n_vol):
                    lab = fn()
                    if drain is not None and k == n_vol - 1:
                        drain()
                    evs[k + 1].record()
                gc.enable()
                barrier()
                ms = evs[0].elapsed_time(evs[-1])
                per_vol.append([round(evs[k].elapsed_time(evs[k + 1]), 2) for k in range(n_vol)])
                if world > 1:
                    t = torch.tensor([ms], device=dev)
                    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                    ms = float(t)
                return ms / n_vol, lab

            n_vol = 8
            per_vol = []
            with torch.no_grad():
                ms_vol, lab = timed_vols(infer_dev, n_vol)
                ms_vol_e2e, lab = timed_vols(infer_e2e, n_vol, e2e_drain if world == 1 else None)
                # post-processing on the device (train.py:167-182; the reference: D2H + scipy on one core + H2D), timed on
                # the label map this (untrained) model predicts -- usually one huge component, the worst case for
                # connected components -- and on a realistic one (three lesions, ~1 % foreground)
                def time_pp(m):
                    """median over 7 calls, each bracketed by its own events (a host stall between two of the ~21
                    launches of a call -- allocator, GC -- would otherwise be charged to the device)"""
                    for _ in range(4):
                        post_process_segment(m, 50)
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(7):
                        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        p0.record()
                        out_m, _ = post_process_segment(m, 50)
                        p1.record()
                        torch.cuda.synchronize()
                        ts.append(p0.elapsed_time(p1))
                    return sorted(ts)[len(ts) // 2], out_m
                pp_ms, pp_mask = time_pp(lab[0, 0])
                _, lesions = synthetic.make_batch(1, 2, tuple(vshape[2:]), seed=5)
                lesions = lesions[0, 0].to(dev)
                pp_ms_real, pp_mask_real = time_pp(lesions)
                # the reference's validation loop for one subject through the reference-facing API, with ITS settings
                # (train.py:148-234: sw_batch_size 2, overlap 0.25 -> 9 forwards of 2 windows; loss; softmax >= 0.5;
                # post-processing; confusion counts), everything on the device; one read of the metrics at the end
                val = None
                if world == 1:
                    try:
                        from fcd_b200 import evaluation, metrics as fmetrics
                        lab_vol = synthetic.make_batch(1, 2, tuple(vshape[2:]), seed=5)[1].to(dev)
                        def time_val(bs):
                            for _ in range(2):
                                evaluation.evaluate_subject(model, vol_d, lab_vol, params, loss_fn, sw_batch_size=bs)
                            torch.cuda.synchronize()
                            acc = fmetrics.VoxelMetricAccumulator()
                            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            n_val = 4
                            v0.record()
                            for _ in range(n_val):
                                vl, vp, vt = evaluation.evaluate_subject(model, vol_d, lab_vol, params, loss_fn,
                                                                         sw_batch_size=bs)
                                acc.update(vp, vt)
                            vm = acc.aggregate()
                            v1.record()
                            torch.cuda.synchronize()
                            return v0.elapsed_time(v1) / n_val, float(vl), vm
                        ms_ref, vl, vm = time_val(2)
                        ms_full, vl2, vm2 = time_val(18)
                        val = {"what": "fcd_b200.evaluate_subject per 256x256x192 subject: sliding window (overlap 0.25) + loss "
                                       "+ label map + device post-processing + confusion counts; metrics read once at the "
                                       "end.  sw_batch_size 2 = the reference's hard-coded value (9 forwards of 2 windows), "
                                       "18 = all windows in one forward (same arithmetic; the deep levels pick their split-K "
                                       "order from the row count, so logits agree to bf16 rounding and labels differ on "
                                       "near-ties only -- an untrained model has many)",
                               "ms_per_subject": ms_ref, "subjects_per_s": 1e3 / ms_ref,
                               "ms_per_subject_sw_batch_18": ms_full, "subjects_per_s_sw_batch_18": 1e3 / ms_full,
                               "val_loss": vl, "metrics": {k: (None if v != v else v) for k, v in vm.items()},
                               "val_loss_sw_batch_18": vl2,
                               "metrics_sw_batch_18": {k: (None if v != v else v) for k, v in vm2.items()}}
                    except Exception as e:      # an extra: never lose the volume numbers over it
                        val = {"error": f"{type(e).__name__}: {e}"[:200]}
            errs = _kernel_errors()
            bad = torch.tensor([float(errs["word"] != 0)], device=dev)
            if world > 1:                      # a time-out on ANY rank voids the volume every rank contributed to
                torch.distributed.all_reduce(bad, op=torch.distributed.ReduceOp.MAX)
            if float(bad) > 0:
                raise RuntimeError(f"tcgen05 pipeline time-out during inference (this rank: {errs})")
            fwd_flops = sum(v["flops"] for k, v in agg.items() if k.endswith((":fwd", ":deconv_fwd"))) / B   # per window
            aux = {"metric": "sliding_window_vols_per_s", "value": 1e3 / ms_vol, "unit": "vols/s", "n_gpus": world,
                   "ms_per_vol": ms_vol, "ms_each_vol_rank0": per_vol, "scaling": "strong", "dtype": "bf16",
                   "config": {"workload": f"{args.model} eval, 2ch 256x256x192 volume, roi {args.patch}^3, overlap 0.5, 18 windows"
                                          + (f" dealt over {world} ranks, reduce-scatter + slab finalize + label all-gather"
                                             if world > 1 else " in one forward") + ", argmax uint8 label map"},
                   "e2e": {"value": 1e3 / ms_vol_e2e, "unit": "vols/s", "ms_per_vol": ms_vol_e2e,
                           "h2d_bytes_per_step": vol.numel() * 4, "d2h_bytes_per_step": lab_h.numel() * world,
                           "pipeline": ("H2D of volume k+1 and D2H of label map k-1 on a copy stream while volume k computes"
                                        if world == 1 else "serial: H2D on rank 0, NVLink broadcast, compute, D2H")},
                   "roofline": {"bound": "tensor", "algorithmic_tflop_per_vol": 18 * fwd_flops / 1e12,
                                "achieved": 18 * fwd_flops / (ms_vol * 1e-3) / 1e12 / world, "unit": "TFLOP/s per GPU",
                                "peak": pk["tf_sust"], "frac": 18 * fwd_flops / (ms_vol * 1e-3) / 1e12 / world / pk["tf_sust"]},
                   "post_process": {"what": "fcd_post_process on the device (opening, 5^3 fill-holes, 26-connected components, "
                                            "size filter 50); not part of value / e2e",
                                    "ms_synthetic_lesions": pp_ms_real, "fg_fraction_synthetic": float(lesions.mean()),
                                    "kept_voxels_synthetic": int(pp_mask_real.sum().item()),
                                    "ms_model_label_map": pp_ms, "fg_fraction_model": float(lab_h.float().mean()),
                                    "kept_voxels_model": int(pp_mask.sum().item())},
                   "validation_loop": val}
        except Exception as e:
            aux = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        return
    cpu = None
    gpu_ref = None
    ar_desc = ("2 buckets, early bucket overlapped with backward" if (reducer.overlap and reducer.early_launches)
               else "1 bucket after backward")
    if world == 1 and not args.no_gpu_reference:
        del model, opt, reducer
        torch.cuda.empty_cache()
        gpu_ref = gpu_reference(args.model, args.patch, B, loss_over, dev)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pps, sec = cpu_reference(args.model, args.patch, args.cpu_sample, warm=1, loss_over=loss_over)
        cpu = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_sample} patches of batch 1 (fwd + {loss_over['loss']} + bwd) after 1 warm-up, fp32, "
                         f"{cores} host threads; ~{sec:.1f} s per patch"}
        if aux and "error" not in aux:
            aux["cpu_baseline"] = cpu_sliding_window(args.model, args.patch, cores)
    nbytes_in = sx.numel() * 4 + sy.numel() * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} train step: fwd + {loss_over['loss']}"
                               f"{' + TV ' + str(args.tv) if args.tv > 0 else ''} + bwd + AdamW, 2ch {args.patch}^3 patches, "
                               f"batch {B}/GPU, dropout p=0.1 active", "global_batch": B * world,
                   "optimizer": "torch.optim.AdamW(fused=True)" if args.torch_adamw else "fcd_b200.optim.FusedAdamW (1 launch)",
                   "parallelism": f"dp{world}", "cuda_graph": graph_ok, "l2": "inputs_exceed_l2",
                   **({"grad_allreduce": ar_desc} if world > 1 else {})},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": gpu_launches * args.steps, "gpu_launches_per_step": gpu_launches,
        "roofline": roof, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "clocks": clocks, "aux": aux,
        "final_loss": final_loss,
    }
    if graph_err:
        line["config"]["graph_error"] = graph_err
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
