"""Calibration of the bf16 parity tolerances: how far is STOCK PyTorch (cuDNN/cuBLAS eager) under bf16 / fp16 autocast
from its own fp32 run on the golden cases?  Run on the GPU box: python tools/calibrate_bf16.py [case ...]

Prints, per case: logits relative L2, argmax flips, loss, and the per-parameter gradient error (worst and
parameter-count-weighted mean) of the autocast run against the fp32 run.  tests/test_gpu_models.py quotes these.
"""
import sys

import torch

sys.path.insert(0, ".")
from oracle import losses as olosses  # noqa: E402
from oracle import nets as onets  # noqa: E402
from tests import helpers as H  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(meta, dtype):
    sd, x, y, noise = H.case_inputs(meta)
    fk = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    leaf = {k: (v.cuda().clone().requires_grad_(k in fk)) for k, v in sd.items()}
    ctx = torch.autocast("cuda", dtype=dtype) if dtype is not None else torch.autocast("cuda", enabled=False)
    with ctx:
        out = onets.forward(meta["model_type"], leaf, x.cuda(), True, {}, noise.cuda())
        vae = None
        if isinstance(out, tuple):
            out, vae = out
        p = H.loss_params(meta)
        loss = olosses.combined_loss(p, out, y.cuda()) if dtype is None else _loss_cuda(p, out, y.cuda())
        total = loss + (p["loss_vae_weight"] * vae if vae is not None else 0.0)
    total.backward()
    return out.detach().float(), float(loss.detach()), {k: leaf[k].grad for k in fk}


def _loss_cuda(p, out, y):
    # oracle loss builds its CE weight on the CPU; move it
    import torch.nn.functional as F
    d = olosses.dice(out, y, squared_pred=p.get("square_pred", False), jaccard=p.get("jaccard", False))
    kind = p.get("loss", "DiceLoss")
    if kind == "DiceCELoss":
        w = torch.tensor([p["ce_background_weight"], p["ce_fcd_weight"]], device=out.device)
        d = p["lambda_dice"] * d + p["lambda_ce"] * F.cross_entropy(out.float(), y.squeeze(1).long(), weight=w)
    elif kind == "DiceFocalLoss":
        d = p["lambda_dice"] * d + p["lambda_focal"] * olosses.focal(out, y, p["gamma_focal"])
    return d


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


if __name__ == "__main__":
    cases = sys.argv[1:] or ["baseunet_p64", "ms_dsa_net_p64", "segresnet_p32"]
    olosses_cross = olosses.cross_entropy
    olosses.cross_entropy = lambda pred, target, wb=0.5, wf=0.5: torch.nn.functional.cross_entropy(
        pred.float(), target.squeeze(1).long(), weight=torch.tensor([wb, wf], device=pred.device))
    for name in cases:
        meta, _ = H.load_case(name)
        meta = dict(meta)
        meta["loss_params"] = {k: v for k, v in meta["loss_params"].items() if not k.startswith("tv")}
        ref, lref, gref = run(meta, None)
        for dt in (torch.bfloat16, torch.float16):
            out, l, g = run(meta, dt)
            worst, ws, ns = ("", 0.0), 0.0, 0
            for k, v in g.items():
                if v is None or gref[k] is None or float(gref[k].norm()) < 1e-7 * gref[k].numel() ** 0.5:
                    continue
                e = rel(v, gref[k])
                ws += e * v.numel()
                ns += v.numel()
                if e > worst[1]:
                    worst = (k, e)
            print(f"{name} {str(dt):15s} logits rel {rel(out, ref):.3e} flips "
                  f"{float((out.argmax(1) != ref.argmax(1)).float().mean()):.4f} loss {l:.5f} vs {lref:.5f} "
                  f"grads weighted-mean {ws / max(ns, 1):.3e} worst {worst[0]} {worst[1]:.3e}")
