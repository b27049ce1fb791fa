"""2+ ranks: three CUDA-graph training steps with the overlapped two-bucket gradient all-reduce against the same steps
with one all-reduce after backward; parameters must end bit-identical, and identical across ranks.
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_overlap.py"""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import parallel, synthetic

rank, local, world = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)


def run(overlap, patch=64):
    params = fcd_b200.get_default_params()
    params.update(model_type="ms_dsa_net", patch_size=(patch,) * 3, loss="DiceCELoss", dropout_rate=0.0)
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    model.apply(synthetic.initialize_weights)
    model = model.to(dev).train()
    for m in model.modules():                      # no dropout: the two runs must see the same arithmetic
        if isinstance(m, torch.nn.Dropout) or isinstance(m, torch.nn.Dropout3d):
            m.p = 0.0
    loss_fn = fcd_b200.CombinedLoss(params, dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    red = parallel.GradAllReducer(model.parameters(), overlap=overlap)
    red.sync_params()
    x, y = synthetic.make_batch(2, 2, patch, seed=10 + rank, device=dev)

    def fwd_bwd():
        loss = loss_fn(model(x), y)
        loss.backward()
        return loss.detach()

    for _ in range(2):                             # eager: observe + first overlapped step
        opt.zero_grad(set_to_none=True)
        fwd_bwd()
        red.allreduce()
        opt.step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.zero_grad(set_to_none=True)
        fwd_bwd()
        red.allreduce()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    opt.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd_bwd()
    red.allreduce(early_in_graph=red.early_captured)
    opt.step()
    for _ in range(3):
        g.replay()
        red.allreduce(early_in_graph=red.early_captured)
        opt.step()
    torch.cuda.synchronize()
    return torch.cat([p.detach().flatten() for p in model.parameters()]), red


a, ra = run(True)
b, rb = run(False)
same = torch.equal(a, b)
gathered = [torch.empty_like(a) for _ in range(world)]
torch.distributed.all_gather(gathered, a)
across = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    print(f"overlap planned={ra.overlap} early_params={len(ra._early[0]) if ra._early else 0} "
          f"early_numel={ra._early[2] if ra._early else 0}/{ra.flat.numel()} captured={ra.early_captured} "
          f"launches={ra.early_launches}")
    print("overlapped == single-bucket parameters:", same, " max|diff| =", float((a - b).abs().max()))
    print("identical across ranks:", across)
torch.distributed.destroy_process_group()
sys.exit(0 if (same and across) else 1)
