"""Which piece of the sharded sliding-window call stalls: graph replay vs all-reduce.  torchrun tools/sw_stall2.py"""
import contextlib, io, sys, time, os
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import parallel, synthetic, inferers, ops
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
nwin = 5 if rank < 2 else 4
mode = os.environ.get("MODE", "graph")
big = torch.zeros(25 * 2 ** 20, device=dev)
rows = []
with torch.no_grad():
    gf = inferers._GraphedWindowForward.get(model, (nwin, 128, 128, 128, 16), dev)
    for i in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "graph":
            gf.graph.replay()
        else:
            model.forward_cl(gf.x)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dist.all_reduce(big)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
from fcd_b200 import _lib
L = _lib.lib()
errs = {n: getattr(L, n)() for n in ("fcd_tc_error", "fcd_tcf_error", "fcd_gemm_tc_error", "fcd_wgrad_tc_error", "fcd_wgrad_gemm_tc_error")}
print(f"rank {rank} errors {errs}", flush=True)
print(f"rank {rank} [{mode}]: " + "  ".join(f"{a:.1f}/{b:.1f}" for a, b in rows), flush=True)
dist.destroy_process_group()
