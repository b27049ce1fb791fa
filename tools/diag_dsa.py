"""DSA-only forward/backward against the oracle on the GPU (development diagnostic)."""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops
from fcd_b200.networks.blocks import TransformerBlock
from oracle import nets as onets, synth
import torch.nn.functional as F

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

for C, dims, P in [(32, (4, 4, 4), 64), (8, (4, 6, 4), 64), (64, (2, 4, 4), 32)]:
    N = dims[0] * dims[1] * dims[2]
    blk = TransformerBlock(input_size=N, hidden_size=C, proj_size=P, num_heads=4, dropout_rate=0.0, pos_embed=True)
    sd = synth.synthetic_state_dict(synth.spec_of(blk.state_dict()), seed=7)
    blk.load_state_dict(sd)
    blk = blk.to(dev).train()
    B = 2
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, *dims, generator=g).to(torch.bfloat16).float().to(dev)
    sdg = {("b." + k): v.to(dev).clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    t = xr.reshape(B, C, N).permute(0, 2, 1) + sdg["b.pos_embed"]
    ln = F.layer_norm(t, (C,), sdg["b.norm.weight"], sdg["b.norm.bias"], 1e-5)
    ref = t + sdg["b.gamma"] * onets.dsa(sdg, "b.dsa", ln)          # [B,N,C]
    dy = torch.randn(B, N, C, generator=g).to(torch.bfloat16).float().to(dev)
    names = ["b.pos_embed", "b.norm.weight", "b.norm.bias", "b.gamma", "b.dsa.qkvv.weight", "b.dsa.EF",
             "b.dsa.temperature", "b.dsa.temperature2"]
    grads = torch.autograd.grad(ref, [xr] + [sdg[k] for k in names], dy)
    xc = ops.to_channels_last(x).requires_grad_(True)
    tt, lnn = ops.ln_pos(xc, blk.pos_embed, blk.norm.weight, blk.norm.bias, C, 1e-5)
    y = blk.dsa(lnn, tt, blk.gamma)
    got = y[..., :C].reshape(B, N, C).float()
    print(f"C={C} N={N} P={P}: fwd rel {rel(got, ref):.3e}  ln rel {rel(lnn[..., :C].reshape(B, N, C).float(), ln):.3e}")
    dyc = torch.zeros_like(y)
    dyc[..., :C] = dy.reshape(B, *dims, C).to(torch.bfloat16)
    y.backward(dyc)
    print("   dx", f"{rel(xc.grad[..., :C].reshape(B, N, C).float(), grads[0].reshape(B, C, N).permute(0, 2, 1)):.3e}")
    mine = dict(blk.named_parameters())
    for k, gr in zip(names, grads[1:]):
        print("  ", k, f"{rel(mine[k[2:]].grad, gr):.3e}", f"|ref|={float(gr.norm()):.3e}")
