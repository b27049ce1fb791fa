"""One eager training step of the bench workload bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (launch list / full capture).  python tools/profile_step.py [model] [patch] [batch]"""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import synthetic

model_type = sys.argv[1] if len(sys.argv) > 1 else "ms_dsa_net"
patch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
params = fcd_b200.get_default_params()
params.update(model_type=model_type, patch_size=(patch,) * 3, loss="DiceCELoss")
torch.manual_seed(42)
with contextlib.redirect_stdout(io.StringIO()):
    model, params = fcd_b200.get_model(params)
model.apply(synthetic.initialize_weights)
model = model.to(dev).train()
loss_fn = fcd_b200.CombinedLoss(params, dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
x, y = synthetic.make_batch(batch, 2, patch, seed=0, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    out = model(x)
    if isinstance(out, tuple):
        out = out[0]
    loss = loss_fn(out, y)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
l = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(l))
