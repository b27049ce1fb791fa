"""Per-entry-point breakdown of one eval-mode window forward (batch 2, 128^3) and of a whole sliding-window volume.
python tools/infer_breakdown.py"""
import contextlib, io, sys, time
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import synthetic, _lib, ops
from fcd_b200.inferers import sliding_window_inference

dev = torch.device("cuda:0")
params = fcd_b200.get_default_params()
params.update(model_type="ms_dsa_net", patch_size=(128,) * 3, loss="DiceCELoss")
torch.manual_seed(42)
with contextlib.redirect_stdout(io.StringIO()):
    model, params = fcd_b200.get_model(params)
model.apply(synthetic.initialize_weights)
model = model.to(dev).eval()
x = torch.randn(2, 128, 128, 128, 16, device=dev).to(torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        model.forward_cl(x)
    torch.cuda.synchronize()
    prof = _lib.Profiler(); _lib.set_profiler(prof)
    model.forward_cl(x)
    agg = prof.summary(); _lib.set_profiler(None)
    tot = sum(a["ms"] for a in agg.values())
    print(f"eval forward (batch 2): sum of library calls {tot:.2f} ms, launches {prof.launches}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:14]:
        print(f"{a['ms']:8.3f} ms  {a['calls']:4d} calls  {k}")
    vol = torch.randn((1, 2, 256, 256, 192)).pin_memory()
    vol_d = torch.empty_like(vol, device=dev)
    for _ in range(2):
        vol_d.copy_(vol, non_blocking=True)
        sliding_window_inference(vol_d, 128, 2, model, overlap=0.5, label_mode="argmax")
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    vol_d.copy_(vol, non_blocking=True)
    ev[1].record()
    _, lab = sliding_window_inference(vol_d, 128, 2, model, overlap=0.5, label_mode="argmax")
    ev[2].record()
    lab_h = lab.to("cpu")
    ev[3].record()
    torch.cuda.synchronize()
    print(f"volume: H2D {ev[0].elapsed_time(ev[1]):.2f} ms, sliding window {ev[1].elapsed_time(ev[2]):.2f} ms, "
          f"D2H labels {ev[2].elapsed_time(ev[3]):.2f} ms")
    g = list(fcd_b200.inferers._GraphedWindowForward._cache.values())[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(9):
        g.graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"9 graph replays of the window forward: {e0.elapsed_time(e1):.2f} ms")
    lab_pin = torch.empty((1, 1, 256, 256, 192), dtype=torch.uint8).pin_memory()
    for bs in (2, 3, 6, 9, 18):
        for _ in range(2):
            sliding_window_inference(vol_d, 128, bs, model, overlap=0.5, label_mode="argmax")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            vol_d.copy_(vol, non_blocking=True)
            _, lab = sliding_window_inference(vol_d, 128, bs, model, overlap=0.5, label_mode="argmax")
            lab_pin.copy_(lab, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print(f"sw_batch_size {bs:2d}: {e0.elapsed_time(e1) / 3:.2f} ms per volume (H2D + windows + D2H to pinned)")
