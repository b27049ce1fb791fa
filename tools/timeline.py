"""Kernel timeline of one CUDA-graph replay of the bench training step (torch.profiler / CUPTI activity records):
per-stream busy time, gaps on the main stream and what they were waiting for, the tail of the step.
python tools/timeline.py [model] [patch] [batch] > gpurun_out/timeline.txt"""
import contextlib, io, json, sys, collections, re
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


def fwd_bwd():
    out = model(x)
    vae = None
    if isinstance(out, (tuple, list)):
        out, vae = out
    loss = loss_fn(out, y)
    if vae is not None:
        loss = loss + params["loss_vae_weight"] * vae
    loss.backward()
    return loss.detach()


for _ in range(2):
    opt.zero_grad(set_to_none=True)
    fwd_bwd()
    opt.step()
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
with torch.cuda.graph(graph):
    loss = fwd_bwd()
for _ in range(3):
    graph.replay()
    opt.step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        graph.replay()
        opt.step()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/trace.json")
ev = json.load(open("gpurun_out/trace.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ks.sort(key=lambda e: e["ts"])
print("kernels recorded", len(ks))
# second replay = second half
half = len(ks) // 2
ks = ks[half:]
t0 = ks[0]["ts"]
def nm(e):
    n = e["name"]
    n = n.replace("(anonymous namespace)::", "").replace("void ", "")
    n = re.sub(r"\(.*", "", n)
    return re.sub(r"<.*", "", n)[:40]
end = max(e["ts"] + e["dur"] for e in ks)
print(f"step span {end - t0:.0f} us, kernels {len(ks)}")
bys = collections.defaultdict(float)
for e in ks:
    bys[e["args"].get("stream")] += e["dur"]
for s, t in sorted(bys.items(), key=lambda x: -x[1]):
    print(f"stream {s}: busy {t:.0f} us")
# concurrency profile: time with k kernels in flight
pts = []
for e in ks:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
conc = collections.defaultdict(float); cur = 0; last = pts[0][0]
for t, d in pts:
    conc[cur] += t - last; last = t; cur += d
print("time (us) with k kernels in flight:", {k: round(v) for k, v in sorted(conc.items())})
# full timeline, compact
print("\n# timeline (start us, dur us, stream, kernel, grid)")
for e in ks:
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} s{e['args'].get('stream')} {nm(e)} g{e['args'].get('grid')}")
