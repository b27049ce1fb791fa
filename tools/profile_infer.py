"""One sliding-window volume (configs[4]: 256x256x192, roi 128^3, overlap 0.5, 18 windows) + the device post-processing,
bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off -k regex:sw_|pp_|uf_|cc_ ...`."""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import synthetic
from fcd_b200.inferers import post_process_segment, sliding_window_inference

dev = torch.device("cuda:0")
params = fcd_b200.get_default_params()
params.update(model_type="ms_dsa_net", patch_size=(128,) * 3)
torch.manual_seed(42)
with contextlib.redirect_stdout(io.StringIO()):
    model, params = fcd_b200.get_model(params)
model.apply(synthetic.initialize_weights)
model = model.to(dev).eval()
vol = torch.randn((1, 2, 256, 256, 192), generator=torch.Generator().manual_seed(7)).to(dev)
_, mask = synthetic.make_batch(1, 2, (256, 256, 192), seed=5)
mask = mask[0, 0].to(dev)                      # a realistic (~1 % foreground, 3 blobs) prediction for the post-processing
with torch.no_grad():
    for _ in range(2):
        _, lab = sliding_window_inference(vol, 128, 18, model, overlap=0.5, label_mode="argmax", return_logits=False)
        post_process_segment(mask, 50)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    _, lab = sliding_window_inference(vol, 128, 18, model, overlap=0.5, label_mode="argmax", return_logits=False)
    m, l = post_process_segment(mask, 50)
    m2, l2 = post_process_segment(lab[0, 0], 50)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("fg fraction", float(lab.float().mean()), "kept", float(m.sum()), float(m2.sum()))
