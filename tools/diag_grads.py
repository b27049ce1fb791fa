"""Per-parameter gradient error of a golden model case (GPU) -- development diagnostic."""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
from tests import helpers as H
from tests.test_gpu_models import build, rel
import fcd_b200

name = sys.argv[1] if len(sys.argv) > 1 else "baseunet_p64"
meta, z = H.load_case(name)
model, params, sd, x, y, noise = build(meta)
ora = H.oracle_run(meta, training=True)
model.train()
if hasattr(model, "set_vae_noise"):
    model.set_vae_noise(noise.cuda())
out = model(x.cuda())
vae = None
if isinstance(out, tuple):
    out, vae = out
loss = fcd_b200.CombinedLoss(params, "cuda")(out, y.cuda())
total = loss + (params["loss_vae_weight"] * vae if vae is not None else 0.0)
total.backward()
print("logits rel", rel(out.detach().cpu(), ora["logits"]), "loss", float(loss), ora["loss"])
for k, p in model.named_parameters():
    og = ora["grads"].get(k)
    if og is None or p.grad is None:
        print(f"{k:60s} none")
        continue
    print(f"{k:60s} rel={rel(p.grad.cpu(), og):9.3e} |og|={float(og.norm()):9.3e} |g|={float(p.grad.norm()):9.3e} n={p.numel()}")
