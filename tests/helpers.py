"""Shared helpers for the parity tests: golden loading and oracle evaluation of a golden model case."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import losses as olosses
from oracle import nets as onets
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODEL_CASES = ["baseunet_p64", "ms_dsa_net_p64", "ms_dsa_net_ps_p64", "segresnet_p32", "segresnetvae_p32",
               "segresnet_dsa_p32", "segresnetvae_dsa_p32"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def case_inputs(meta):
    spec = [(k, tuple(s), d) for k, s, d in meta["spec"]]
    sd = synth.synthetic_state_dict(spec, seed=meta["weights_seed"])
    x = synth.image(meta["batch"], 2, meta["patch"], seed=meta["image_seed"])
    y = synth.label(meta["batch"], meta["patch"], seed=meta["label_seed"])
    noise = synth.tensor((meta["batch"], 256), "vae_noise", meta["noise_seed"], 1.0, dist="normal")
    return sd, x, y, noise


def loss_params(meta):
    p = dict(loss="DiceLoss", lambda_dice=1.0, lambda_ce=1.0, lambda_focal=1.0, ce_background_weight=0.5,
             ce_fcd_weight=0.5, gamma_focal=2.0, jaccard=False, square_pred=False, sigmoid=False, softmax=True,
             tv_loss_norm="l1", tv_loss_weight=0.0, tvloss_exclude_borders=False, chans_out=2,
             loss_vae_weight=0.2)
    p.update(meta["loss_params"])
    return p


def oracle_run(meta, training=True, device="cpu", autocast=None):
    """Oracle forward (+loss +grads when training).  Returns dict(logits, loss, total, grads, bn).

    device/autocast: the SAME functional oracle evaluated by stock PyTorch on the GPU under bf16/fp16 autocast is
    what calibrates the bf16 tolerances of the GPU parity tests (see tools/calibrate_bf16.py)."""
    import contextlib
    sd, x, y, noise = case_inputs(meta)
    float_keys = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    leaf = {k: (v.to(device).clone().requires_grad_(training) if k in float_keys else v.to(device))
            for k, v in sd.items()}
    x, y, noise = x.to(device), y.to(device), noise.to(device)
    bn = {}
    ctx = torch.autocast("cuda", dtype=autocast) if autocast is not None else contextlib.nullcontext()
    with ctx:
        out = onets.forward(meta["model_type"], leaf, x, training, bn, noise)
        vae = None
        if isinstance(out, tuple):
            out, vae = out
        res = dict(logits=out.detach().float().cpu(), bn={k: v.detach().cpu() for k, v in bn.items()})
        if training:
            p = loss_params(meta)
            loss = olosses.combined_loss(p, out, y)
            total = loss + (p["loss_vae_weight"] * vae if vae is not None else 0.0)
    if training:
        total.backward()
        res.update(loss=float(loss.detach()), total=float(total.detach()),
                   vae_loss=None if vae is None else float(vae.detach()),
                   grads={k: (None if leaf[k].grad is None else leaf[k].grad.detach().float().cpu())
                          for k in float_keys})
    return res


def probe(name, t):
    v = synth.tensor(t.shape, "probe:" + name, 0, 1.0)
    return float(t.double().norm()), float((t.double() * v.double()).sum())
