"""torch.matmul (cuBLAS) on the GEMM shapes of the bench workload - run under ncu to see which kernels / grids / clusters the
library picks (tools/README.md), or alone for timings. Not part of the product path."""
import sys
import torch

shapes = [(2414, 12288, 4096), (2414, 4096, 4096), (2414, 24576, 4096), (2414, 4096, 12288), (4682, 12288, 4096)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        c = a @ w.t()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        c = a @ w.t()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas {M}x{N}x{K}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
