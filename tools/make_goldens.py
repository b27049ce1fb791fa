#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own source files (from /root/reference, through
oracle/monai_shim.py for the absent MONAI package) on portable synthetic weights / inputs (oracle/synth.py).

Run in the build container only:   python tools/make_goldens.py
The GPU box has no /root/reference; it consumes the committed .npz files.
"""
from __future__ import annotations

import json
import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = [
    # name, model_type, patch, feature_size, batch, loss overrides
    ("baseunet_p64", "baseunet", 64, 4, 1, dict(loss="DiceCELoss")),
    ("ms_dsa_net_p64", "ms_dsa_net", 64, 4, 2, dict(loss="DiceCELoss")),
    ("ms_dsa_net_ps_p64", "ms_dsa_net_ps", 64, 4, 1, dict(loss="DiceLoss")),
    ("segresnet_p32", "segresnet", 32, 8, 2, dict(loss="DiceFocalLoss")),
    ("segresnetvae_p32", "segresnetvae", 32, 8, 2, dict(loss="DiceFocalLoss")),
    ("segresnet_dsa_p32", "segresnet_dsa", 32, 8, 2, dict(loss="DiceCELoss", tv_loss_weight=0.1, tv_loss_norm="l1")),
    ("segresnetvae_dsa_p32", "segresnetvae_dsa", 32, 8, 1,
     dict(loss="DiceCELoss", tv_loss_weight=0.1, tv_loss_norm="l2", tvloss_exclude_borders=True)),
]


def probe(name, t):
    """Two scalars per tensor: L2 norm and a dot product with a portable probe vector."""
    v = synth.tensor(t.shape, "probe:" + name, 0, 1.0)
    return float(t.double().norm()), float((t.double() * v.double()).sum())


def run_model_case(name, model_type, patch, fs, batch, over):
    _, gl, _ = ref_loader.load()
    params = ref_loader.default_params()
    params.update(model_type=model_type, patch_size=(patch,) * 3, feature_size=fs)
    params.update(over)
    model, params = ref_loader.build_model(params)
    spec = synth.spec_of(model.state_dict())
    sd = synth.synthetic_state_dict(spec, seed=1)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    model.train()
    x = synth.image(batch, 2, patch, seed=3)
    y = synth.label(batch, patch, seed=5)
    noise = synth.tensor((batch, 256), "vae_noise", 9, 1.0, dist="normal")
    loss_fn = gl.CombinedLoss(params, torch.device("cpu"))
    with mock.patch.object(torch, "randn_like", lambda t, **k: noise.to(t)):
        out = model(x)
    vae_loss = None
    if isinstance(out, tuple):
        out, vae_loss = out
    loss = loss_fn(out, y)
    total = loss + (params["loss_vae_weight"] * vae_loss if vae_loss is not None else 0.0)
    total.backward()
    grads = {k: (probe(k, p.grad) if p.grad is not None else None) for k, p in model.named_parameters()}
    bn = {k: v.detach().double().numpy() for k, v in model.state_dict().items()
          if k.endswith(("running_mean", "running_var"))}
    bn_probe = {k: [float(np.linalg.norm(v)), float(v.sum())] for k, v in bn.items()}
    model.eval()
    with torch.no_grad():
        out_eval = model(x)
    if isinstance(out_eval, tuple):
        out_eval = out_eval[0]
    meta = dict(model_type=model_type, patch=patch, feature_size=fs, batch=batch, loss_params=over,
                spec=[[k, list(s), d] for k, s, d in spec], weights_seed=1, image_seed=3, label_seed=5,
                noise_seed=9, grads=grads, bn=bn_probe, loss=float(loss.detach()), total=float(total.detach()),
                vae_loss=None if vae_loss is None else float(vae_loss.detach()),
                logits_norm=float(out.double().norm()), logits_sum=float(out.double().sum()),
                eval_logits_norm=float(out_eval.double().norm()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta),
                        logits_sub=out.detach()[:, :, ::3, ::3, ::3].numpy(),
                        eval_logits_sub=out_eval[:, :, ::3, ::3, ::3].numpy(),
                        argmax_sub=out.detach().argmax(1)[:, ::2, ::2, ::2].numpy().astype(np.uint8))
    print(name, "loss", float(loss.detach()), "total", float(total.detach()), "params", len(spec))


def run_loss_cases():
    _, gl, _ = ref_loader.load()
    base = ref_loader.default_params()
    res = {}
    cfgs = {
        "dice": dict(loss="DiceLoss"),
        "dice_sq_jac": dict(loss="DiceLoss", square_pred=True, jaccard=True),
        "dicece": dict(loss="DiceCELoss"),
        "dicece_w": dict(loss="DiceCELoss", ce_background_weight=0.3, ce_fcd_weight=0.7, lambda_ce=0.5),
        "dicefocal": dict(loss="DiceFocalLoss"),
        "dicefocal_g3": dict(loss="DiceFocalLoss", gamma_focal=3.0, lambda_focal=2.0),
        "dicece_tv_l1": dict(loss="DiceCELoss", tv_loss_weight=0.1),
        "dicece_tv_l2": dict(loss="DiceCELoss", tv_loss_weight=0.1, tv_loss_norm="l2"),
        "dicece_tv_l1_xb": dict(loss="DiceCELoss", tv_loss_weight=0.1, tvloss_exclude_borders=True),
        "dicefocal_tv_l2_xb": dict(loss="DiceFocalLoss", tv_loss_weight=0.2, tv_loss_norm="l2",
                                   tvloss_exclude_borders=True),
        "gdice": dict(loss="GeneralizedDiceLoss"),
        "gdice_simple": dict(loss="GeneralizedDiceLoss", gdice_wtype="simple"),
        "gdice_uniform": dict(loss="GeneralizedDiceLoss", gdice_wtype="uniform"),
        "gdicefocal": dict(loss="GeneralizedDiceFocalLoss", lambda_dice=0.7, lambda_focal=2.0, gamma_focal=3.0),
        "gdicefocal_simple_tv": dict(loss="GeneralizedDiceFocalLoss", gdice_wtype="simple", tv_loss_weight=0.1),
    }
    pred = synth.tensor((2, 2, 20, 24, 28), "loss_pred", 11, 2.0, dist="normal")
    tgt = synth.label(2, (20, 24, 28), seed=13)
    for k, over in cfgs.items():
        p = dict(base)
        p.update(over)
        fn = gl.CombinedLoss(p, torch.device("cpu"))
        pr = pred.clone().requires_grad_(True)
        l = fn(pr, tgt)
        l.backward()
        res[k] = dict(params=over, loss=float(l.detach()), grad_norm=float(pr.grad.double().norm()),
                      grad_probe=probe("lossgrad", pr.grad)[1])
        np.save(os.path.join(OUT, f"lossgrad_{k}.npy"), pr.grad[:, :, ::2, ::2, ::2].numpy())
        print("loss", k, float(l.detach()))
    with open(os.path.join(OUT, "loss_cases.json"), "w") as f:
        json.dump(dict(pred_key="loss_pred", pred_seed=11, pred_scale=2.0, shape=[2, 2, 20, 24, 28],
                       label_seed=13, cases=res), f, indent=1)


def run_sliding_and_postproc():
    from monai.inferers import sliding_window_inference as ref_swi  # the shim's restatement of MONAI
    _, _, uc = ref_loader.load()
    w = synth.tensor((2, 2, 3, 3, 3), "sw_w", 0, 0.3)

    def predictor(x):
        return torch.nn.functional.conv3d(x, w, padding=1)

    cases = {}
    for name, size, roi, ov, bs in [("a", (80, 72, 48), 32, 0.25, 2), ("b", (80, 72, 48), 32, 0.5, 4),
                                    ("c", (24, 40, 32), 32, 0.5, 2), ("d", (64, 64, 32), 32, 0.5, 3)]:
        x = synth.image(1, 2, size, seed=17)
        out = ref_swi(inputs=x, roi_size=(roi,) * 3, sw_batch_size=bs, predictor=predictor, overlap=ov)
        cases[name] = dict(size=list(size), roi=roi, overlap=ov, sw_batch_size=bs,
                           norm=float(out.double().norm()), sum=float(out.double().sum()))
        np.save(os.path.join(OUT, f"sw_{name}.npy"), out[:, :, ::4, ::4, ::4].numpy())
    # post-processing: reference's own utils/utils_common.py (numpy + scipy only)
    rng_mask = (synth.tensor((40, 48, 44), "pp_mask", 21, 1.0, dist="normal") > 1.2).numpy()
    blobs = synth.label(1, (40, 48, 44), seed=23, n_blobs=5)[0, 0].numpy() > 0
    mask = (rng_mask | blobs).astype(np.float32)
    pp = {}
    for l_min in (50, 5, -1):
        m, lab = uc.post_process_segment(mask, l_min)
        pp[str(l_min)] = dict(vox=int(m.sum()), ncomp=int(lab.max()), lab_sum=int(lab.sum()))
        np.save(os.path.join(OUT, f"pp_lab_{l_min}.npy"), lab.astype(np.uint8))
    m, lab = uc.post_process_segment(np.zeros((8, 8, 8), np.float32), -1)
    pp["empty_-1"] = dict(vox=int(m.sum()))
    with open(os.path.join(OUT, "sw_pp_cases.json"), "w") as f:
        json.dump(dict(sw=cases, pp=pp), f, indent=1)
    print("sliding/postproc", cases, pp)


GRIDMASK_CASES = [
    # shape, d, (st_d, st_h, st_w), ratio, mode
    ((16, 16, 16), 5, (0, 3, 4), 0.5, 0), ((12, 20, 9), 4, (3, 0, 1), 0.3, 1), ((32, 24, 40), 11, (10, 2, 7), 0.5, 0),
    ((24, 24, 24), 16, (15, 8, 0), 0.5, 0), ((20, 33, 17), 3, (2, 2, 2), 0.9, 1),
]


def run_augment_and_metrics():
    """utils/gridmask.py:8-72 (Grid.__call__ with its np.random draws substituted) and utils/utils_common.py:37-60
    (evaluate_fp) of the reference on small seeded inputs: what oracle/sampling.gridmask and oracle/metrics.evaluate_fp
    are checked against where /root/reference is absent."""
    from scipy import ndimage as nd
    gm = ref_loader.load_gridmask()
    _, _, uc = ref_loader.load()
    masks = {}
    for k, (shape, d, st, ratio, mode) in enumerate(GRIDMASK_CASES):
        draws = iter([d] + list(st) + [0])
        grid = gm.Grid(2, 64, rotate=1, ratio=ratio, mode=mode, prob=1.0)
        with mock.patch.object(np.random, "rand", lambda: 0.0), \
                mock.patch.object(np.random, "randint", lambda *a, **kw: next(draws)):
            out = grid(torch.ones((1,) + tuple(shape)))
        masks[f"mask_{k}"] = np.packbits(out[0].numpy().astype(bool))
    np.savez_compressed(os.path.join(OUT, "gridmask_cases.npz"),
                        meta=json.dumps([dict(shape=list(c[0]), d=c[1], st=list(c[2]), ratio=c[3], mode=c[4])
                                         for c in GRIDMASK_CASES]), **masks)
    fp = []
    for k in range(5):
        pred = nd.binary_dilation(synth.tensor((24, 28, 20), "fp_pred", 40 + k, 1.0, dist="normal").numpy() > 2.6,
                                  iterations=1 + k % 3)
        lab = nd.binary_dilation(synth.tensor((24, 28, 20), "fp_lab", 50 + k, 1.0, dist="normal").numpy() > 2.0,
                                 iterations=2).astype(np.float32)
        cc, n = nd.label(pred)
        fp.append(dict(seed=k, components=int(n), fp=int(uc.evaluate_fp(cc, lab))))
    with open(os.path.join(OUT, "evaluate_fp_cases.json"), "w") as f:
        json.dump(fp, f, indent=1)
    print("gridmask", [int(np.unpackbits(v).sum()) for v in masks.values()], "evaluate_fp", fp)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only in (None, "models"):
        for c in MODEL_CASES:
            run_model_case(*c)
    if only in (None, "loss"):
        run_loss_cases()
    if only in (None, "sliding"):
        run_sliding_and_postproc()
    if only in (None, "augment"):
        run_augment_and_metrics()
