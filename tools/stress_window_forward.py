"""CUDA-graph replays of the whole MS_DSA_NET window forward on ONE GPU, the regime of round 1's open defect.

    FCD_NSEG_UNRESTRICTED=1 python tools/stress_window_forward.py [replays] [windows] [busy]

With `windows` = 5 (what a rank of a 4-GPU sharded inference evaluates) and the chooser unrestricted, the 128^3 convs
run with 2 d-segments (1280 items on 296 CTAs) and the 64^3 / 32^3 levels with 8 short segments, inside a graph whose
branch streams keep other tcgen05 kernels resident on the same SMs.  Every replay must reproduce the first replay's
logits bit for bit and leave the status word at zero; on a time-out the debug record (kernel, wait site, item, the
progress counters of every role of the CTA) is printed.  `busy` > 0 additionally keeps a compute-heavy torch kernel
running on another stream (SM pressure: fewer CTAs of a launch resident at once, as under NCCL).
Prints `RESULT ok` or `RESULT FAILED`."""
import contextlib
import io
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import fcd_b200  # noqa: E402
from fcd_b200 import _lib, inferers, synthetic  # noqa: E402


def main():
    replays = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    busy = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    params = fcd_b200.get_default_params()
    params.update(model_type="ms_dsa_net", patch_size=(128,) * 3)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    model.apply(synthetic.initialize_weights)
    model = model.to(dev).eval()
    L = _lib.lib()
    segs = {lvl: L.fcd_conv3_tc_nseg(nwin, s, s, s, k, n) for lvl, (s, k, n) in
            {"128^3 16->16": (128, 16, 16), "128^3 32->16": (128, 32, 16), "64^3 32->32": (64, 32, 32),
             "64^3 64->32": (64, 64, 32), "32^3 64->32": (32, 64, 32), "32^3 32->32": (32, 32, 32)}.items()}
    print("d-segments chosen:", segs, flush=True)
    bad = 0
    with torch.no_grad():
        gf = inferers._GraphedWindowForward.get(model, (nwin, 128, 128, 128, 16), dev)
        g = torch.Generator().manual_seed(0)
        gf.x.copy_(torch.randn(gf.x.shape, generator=g).to(torch.bfloat16))
        _lib.status()
        gf.graph.replay()
        torch.cuda.synchronize()
        ref = gf.y.clone()
        st = _lib.status()
        if st["word"]:
            print("first replay:", st, flush=True)
            bad += 1
        side = torch.cuda.Stream()
        a = torch.randn(4096, 4096, device=dev)
        times = []
        for i in range(replays):
            if busy:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(busy):
                        a = torch.tanh(a @ a * 1e-4)
            t0 = time.perf_counter()
            if os.environ.get("MODE", "graph") == "eager":      # the same forward launched kernel by kernel
                gf.y.copy_(gf._fwd(model))
            else:
                gf.graph.replay()
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            st = _lib.status()
            same = torch.equal(gf.y, ref)
            if st["word"] or not same:
                bad += 1
                d = (gf.y - ref).abs()
                print(f"replay {i}: {times[-1]:.1f} ms, status {st if st['word'] else 'clean'}, "
                      f"max |dy| {float(d.max()):.4g}, differing voxels {int((d > 0).sum())}", flush=True)
    times.sort()
    print(f"{replays} replays of a {nwin}-window forward: median {times[len(times) // 2]:.2f} ms, max {times[-1]:.2f} ms, "
          f"{bad} bad", flush=True)
    print("RESULT ok" if bad == 0 else "RESULT FAILED", flush=True)
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
