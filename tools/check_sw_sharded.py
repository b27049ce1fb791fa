"""Sharded sliding-window inference on N ranks (one process per GPU, NCCL) against the same call on one rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/check_sw_sharded.py

1. An exact-arithmetic predictor (dyadic conv, tests/test_gpu_inference.py): fp32 sums are exact whatever the
   reduction order, so the sharded results must be BIT-IDENTICAL to the single-rank ones -- both exchange paths:
   all-reduce (logits everywhere) and reduce-scatter + slab finalize + label all-gather (labels only).
2. A real MS_DSA_NET (bf16 kernels): a rank batches other windows together than the single-rank run and NCCL adds the
   partial volumes in another order than the window order, so logits agree to bf16 rounding noise and labels may flip
   only on near-ties.
Prints `RESULT ok` on every rank or raises."""
import contextlib
import io
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcd_b200  # noqa: E402
from fcd_b200 import _lib, parallel, synthetic  # noqa: E402
from fcd_b200.inferers import sliding_window_inference as swi  # noqa: E402
from tests.test_gpu_inference import _ExactNet, _dyadic  # noqa: E402


def main():
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    w = _dyadic((2, 2, 3, 3, 3), 5, 0.125, 1.0)
    net = _ExactNet(w).to(dev).eval()
    cases = [((80, 72, 48), 32, 0.5, 4, 1), ((24, 40, 32), 32, 0.5, 2, 1), ((27, 33, 40), 32, 0.25, 3, 2),
             ((64, 64, 32), 32, 0.5, 5, 1)]
    with torch.no_grad():
        for size, roi, ov, bs, B in cases:
            x = _dyadic((B, 2) + size, 7, 1.0 / 16, 2.0)
            x[:, :, :9, :9, :9] = 0.0
            xd = x.to(dev)
            for lm in ("argmax", "threshold"):
                ref, ref_lab = swi(xd, roi, bs, net, overlap=ov, label_mode=lm)                      # this rank alone
                out, lab = swi(xd, roi, bs, net, overlap=ov, label_mode=lm, shard=True)              # all-reduce path
                assert torch.equal(out, ref) and torch.equal(lab, ref_lab), (size, lm, "all-reduce path")
                none, lab2 = swi(xd, roi, bs, net, overlap=ov, label_mode=lm, shard=True, return_logits=False)
                assert none is None and lab2.dtype == ref_lab.dtype and lab2.shape == ref_lab.shape
                assert torch.equal(lab2, ref_lab), (size, lm, "reduce-scatter path",
                                                    int((lab2 != ref_lab).sum()))
        # real network
        params = fcd_b200.get_default_params()
        params.update(model_type="ms_dsa_net", patch_size=(64,) * 3, feature_size=8)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model, _ = fcd_b200.get_model(params)
        model.apply(synthetic.initialize_weights)
        model = model.to(dev).eval()
        for p in model.parameters():            # identical weights on every rank
            dist.broadcast(p.data, 0)
        for b in model.buffers():
            dist.broadcast(b.data, 0)
        vol = torch.randn((1, 2, 128, 96, 96), generator=torch.Generator().manual_seed(3)).to(dev)
        ref, ref_lab = swi(vol, 64, 4, model, overlap=0.5, label_mode="argmax")
        out, lab = swi(vol, 64, 4, model, overlap=0.5, label_mode="argmax", shard=True)
        _, lab2 = swi(vol, 64, 4, model, overlap=0.5, label_mode="argmax", shard=True, return_logits=False)
        rel = float((out - ref).norm() / ref.norm())
        flips = float((lab != ref_lab).float().mean())
        flips2 = float((lab2 != ref_lab).float().mean())
        print(f"rank {rank}: MS_DSA_NET sharded vs single rank: logits rel L2 {rel:.2e}, label flips {flips:.2e} "
              f"(all-reduce) {flips2:.2e} (reduce-scatter)", flush=True)
        # (the deep levels pick their split-K factor from the rows of a forward call, and a rank batches other windows
        # together than the single-rank run: bf16 rounding noise, as in test_sliding_window_batching_is_invariant)
        assert rel <= 1e-2 and flips <= 5e-3 and flips2 <= 5e-3
        # every rank holds the same label map
        gathered = [torch.empty_like(lab2) for _ in range(world)]
        dist.all_gather(gathered, lab2)
        assert all(torch.equal(g, lab2) for g in gathered)
    _lib.check_errors()
    print(f"rank {rank}: RESULT ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
