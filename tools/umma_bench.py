"""Cycles per tcgen05.mma (M x N x 16, bf16, operands in shared memory, no-swizzle K-major) on this B200.
python tools/umma_bench.py"""
import ctypes
import os
import subprocess
import sys
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libumma_bench.so")
if not os.path.exists(SO):        # measurement aid, deliberately NOT part of the product library
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                           "-Xcompiler", "-fPIC", "-I", os.path.join(HERE, "..", "include"),
                           os.path.join(HERE, "umma_bench.cu"), "-o", SO])
LIB = ctypes.CDLL(SO)
LIB.fcd_umma_bench.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]

out = torch.zeros(1, dtype=torch.int64, device="cuda")
ITERS = 4096


def run(M, N, nissue, same, ctas=1):
    for _ in range(2):
        rc = LIB.fcd_umma_bench(M, N, ITERS, nissue, same, ctas, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    torch.cuda.synchronize()
    return float(out.item()) / (ITERS * nissue)


print("cycles per instruction (per SM), one CTA; dense-rate math would be M*N*16/4096 MAC/clk = N/2 cycles at M=128")
print(f"{'M':>4} {'N':>4} | {'1 warp':>8} {'2 warps':>8} {'3 warps':>8} {'4 warps':>8} | {'3 warps, one accumulator':>25}")
for M in (128, 64):
    for N in (16, 32, 48, 64, 96, 128, 256):
        row = []
        for nw in (1, 2, 3, 4):
            row.append(run(M, N, nw, 0) if N <= 128 else (run(M, N, nw, 1)))
        same = run(M, N, 3, 1)
        print(f"{M:4d} {N:4d} | " + " ".join(f"{v:8.1f}" for v in row) + f" | {same:25.1f}")
print("all 148 SMs busy (one CTA each), M=128, 3 issuing warps:")
for N in (16, 48, 96):
    print(f"  N={N:3d}: {run(128, N, 3, 0, ctas=148):.1f} cycles per instruction")
