"""fcd_b200.optim.FusedAdamW (one fcd_adamw_multi launch for all tensors) against torch.optim.AdamW, the optimizer the
reference builds (train_utils.py:63-71, lr 1e-4, weight_decay 1e-5) -- same arithmetic, same state layout."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(16, 2, 3, 3, 3), (7,), (1,), (33, 5), (512, 512, 3), (4099,), (64, 32, 1, 1, 1), (3, 3)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]


def _grads(params, step):
    g = torch.Generator().manual_seed(100 + step)
    for p in params:
        p.grad = (torch.randn(p.shape, generator=g) * (0.1 + 0.05 * step)).to(DEV)


@pytest.mark.parametrize("cfg", [dict(lr=1e-4, weight_decay=1e-5), dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)])
def test_fused_adamw_matches_torch(cfg):
    from fcd_b200 import _lib
    from fcd_b200.optim import FusedAdamW
    a, b = _params(0), _params(0)
    ours = FusedAdamW(a, **cfg)
    ref = torch.optim.AdamW(b, foreach=False, fused=False, **cfg)
    for step in range(6):
        _grads(a, step)
        _grads(b, step)
        before = _lib.LAUNCHES
        ours.step()
        assert _lib.LAUNCHES == before + 1, "one launch for all parameter tensors"
        ref.step()
        for p, q in zip(a, b):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), (step, p.shape, float((p - q).abs().max()))
    for p, q in zip(a, b):
        assert torch.allclose(ours.state[p]["exp_avg"], ref.state[q]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(ours.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], rtol=1e-5, atol=1e-9)
        assert float(ours.state[p]["step"]) == float(ref.state[q]["step"]) == 6.0


def test_fused_adamw_state_dict_round_trips_with_torch():
    """optimizer_state_dict of a reference checkpoint (train.py:113-146) loads into FusedAdamW and vice versa."""
    from fcd_b200.optim import FusedAdamW
    a, b = _params(1), _params(1)
    ours = FusedAdamW(a, lr=1e-3, weight_decay=1e-2)
    ref = torch.optim.AdamW(b, lr=1e-3, weight_decay=1e-2, foreach=False, fused=False)
    for step in range(3):
        _grads(a, step)
        _grads(b, step)
        ours.step()
        ref.step()
    # swap the states: ours continues from torch's state, torch from ours
    sd_ours, sd_ref = ours.state_dict(), ref.state_dict()
    # torch.optim.AdamW.load_state_dict adopts `step` tensors without copying and then increments each one once per
    # parameter: the state dict must hold one tensor per parameter (the live state shares one device counter)
    steps = [st["step"] for st in sd_ours["state"].values()]
    assert len({id(t) for t in steps}) == len(steps) and all(float(t) == 3.0 and not t.is_cuda for t in steps)
    assert float(ours.state[a[0]]["step"]) == 3.0 and ours.state[a[0]]["step"] is ours.state[a[1]]["step"]
    ours2 = FusedAdamW(a, lr=1e-3, weight_decay=1e-2)
    ours2.load_state_dict(sd_ref)
    ref2 = torch.optim.AdamW(b, lr=1e-3, weight_decay=1e-2, foreach=False, fused=False)
    ref2.load_state_dict(sd_ours)
    for step in range(3, 6):
        _grads(a, step)
        _grads(b, step)
        ours2.step()
        ref2.step()
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=5e-6, atol=1e-7), float((p - q).abs().max())


def test_fused_adamw_inside_cuda_graph_replays():
    """The step counter lives on the device, so a captured step() advances its bias correction on every replay."""
    from fcd_b200.optim import FusedAdamW
    a, b = _params(2), _params(2)
    for p in a + b:
        p.grad = torch.zeros_like(p)
    ours = FusedAdamW(a, lr=1e-2, weight_decay=0.0)
    ref = torch.optim.AdamW(b, lr=1e-2, weight_decay=0.0, foreach=False, fused=False)
    gsrc = [torch.randn_like(p) for p in a]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for p, g in zip(a, gsrc):
            p.grad.copy_(g)
        ours.step()                      # warm-up outside capture (state + job table)
    torch.cuda.current_stream().wait_stream(s)
    for q, g in zip(b, gsrc):
        q.grad.copy_(g)
    ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step()
    for _ in range(3):                   # capture itself does not execute: 3 replays = steps 2..4
        graph.replay()
        ref.step()
    torch.cuda.synchronize()
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=5e-6, atol=1e-7), float((p - q).abs().max())
