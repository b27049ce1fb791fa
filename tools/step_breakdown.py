"""Per-entry-point CUDA-event breakdown of one eager training step (after warm-up).
python tools/step_breakdown.py [model] [patch] [batch]"""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import synthetic, _lib

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


for _ in range(3):
    step()
torch.cuda.synchronize()
prof = _lib.Profiler()
_lib.set_profiler(prof)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
agg = prof.summary()
_lib.set_profiler(None)
tot = sum(a["ms"] for a in agg.values())
print(f"eager step {e0.elapsed_time(e1):.2f} ms; sum of library calls {tot:.2f} ms; launches {prof.launches}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    extra = ""
    if a["flops"]:
        extra = f"  {a['flops'] / a['ms'] / 1e9:8.1f} TF/s"
    elif a["bytes"]:
        extra = f"  {a['bytes'] / a['ms'] / 1e6:8.1f} GB/s"
    print(f"{a['ms']:8.3f} ms  {a['calls']:4d} calls  {k:28s}{extra}")
