"""Diagnose per-call stalls of sharded sliding-window inference: torchrun --nproc-per-node N tools/sw_stall.py"""
import contextlib, io, sys, time
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import parallel, synthetic, inferers
import torch.distributed as dist

rank, local, world = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
params = fcd_b200.get_default_params()
params.update(model_type="ms_dsa_net", patch_size=(128,) * 3)
with contextlib.redirect_stdout(io.StringIO()):
    model, params = fcd_b200.get_model(params)
model.apply(synthetic.initialize_weights)
model = model.to(dev).eval()
vol = torch.randn((1, 2, 256, 256, 192), device=dev)
orig_ar = dist.all_reduce
t_ar = []


def timed_ar(t, *a, **k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = orig_ar(t, *a, **k)
    torch.cuda.synchronize()
    t_ar.append((time.perf_counter() - t0) * 1e3)
    return r


dist.all_reduce = timed_ar
rows = []
with torch.no_grad():
    for i in range(14):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, lab = inferers.sliding_window_inference(vol, 128, 18, model, overlap=0.5, label_mode="argmax")
        torch.cuda.synchronize()
        rows.append(((time.perf_counter() - t0) * 1e3, t_ar[-1] if t_ar else 0.0,
                     torch.cuda.memory_reserved() / 2 ** 20, len(inferers._GraphedWindowForward._cache)))
print(f"rank {rank}: " + "  ".join(f"{a:.1f}/{b:.1f}ms {m:.0f}MB c{c}" for a, b, m, c in rows), flush=True)
dist.destroy_process_group()
