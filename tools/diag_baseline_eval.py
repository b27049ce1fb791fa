"""MS_DSA_NET fs16 128^3 b2: ours vs fp32 CPU oracle vs stock bf16 autocast (the functional oracle on the GPU), train and
eval mode -- separates conditioning (stock is as far off) from a defect (only ours is)."""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
import fcd_b200
from oracle import nets as onets, synth

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


mt = sys.argv[1] if len(sys.argv) > 1 else "ms_dsa_net"
patch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
params = fcd_b200.get_default_params()
params.update(model_type=mt, patch_size=(patch,) * 3, feature_size=16)
with contextlib.redirect_stdout(io.StringIO()):
    model, params = fcd_b200.get_model(params)
sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=1)
model.load_state_dict(sd)
for m in model.modules():
    if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
        m.p = 0.0
model = model.to(DEV)
x = synth.image(batch, 2, patch, seed=3)
sdd = {k: v.to(DEV) for k, v in sd.items()}
for mode in ("train", "eval"):
    tr = mode == "train"
    model.train(tr)
    with torch.no_grad():
        ref = onets.forward(mt, sd, x, tr, {})
        out = model(x.to(DEV)).cpu()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cal = onets.forward(mt, sdd, x.to(DEV), tr, {}).float().cpu()
        with torch.autocast("cuda", dtype=torch.float16):
            cal16 = onets.forward(mt, sdd, x.to(DEV), tr, {}).float().cpu()
        f32 = onets.forward(mt, sdd, x.to(DEV), tr, {}).float().cpu()
    print(f"{mt} {mode}: ours {rel(out, ref):.3e} | stock bf16 autocast {rel(cal, ref):.3e} | stock fp16 autocast "
          f"{rel(cal16, ref):.3e} | stock fp32 GPU {rel(f32, ref):.3e} | logit range {float(ref.min()):.2f}..{float(ref.max()):.2f}",
          flush=True)
