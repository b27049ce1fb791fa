import sys, torch
sys.path.insert(0, '.')
from tests import helpers as H
from oracle import nets as onets, losses as olosses
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for name in ["baseunet_p64", "ms_dsa_net_p64", "segresnet_p32"]:
    meta, z = H.load_case(name)
    sd, x, y, noise = H.case_inputs(meta)
    sdg = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = onets.forward(meta["model_type"], sdg, x.cuda(), True, {}, noise.cuda())
        if isinstance(ref, tuple): ref = ref[0]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            bf = onets.forward(meta["model_type"], sdg, x.cuda(), True, {}, noise.cuda())
        if isinstance(bf, tuple): bf = bf[0]
        with torch.autocast("cuda", dtype=torch.float16):
            hf = onets.forward(meta["model_type"], sdg, x.cuda(), True, {}, noise.cuda())
        if isinstance(hf, tuple): hf = hf[0]
    r = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    print(name, "torch bf16-autocast vs fp32 rel L2:", r(bf.float(), ref), " fp16-autocast:", r(hf.float(), ref),
          " argmax flips bf16:", float((bf.argmax(1) != ref.argmax(1)).float().mean()))
