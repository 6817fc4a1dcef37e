"""ncu target: a few launches of the CTA-pair GEMM kernel and of the 1-CTA kernel on the QKV-sized problem."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmada_parallel_b200 import _lib
torch.manual_seed(0)
M = int(os.environ.get("MMDP_PROF_M", "7242"))
a = (torch.randn(M, 4096, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(12288, 4096, device="cuda") * 0.05).to(torch.bfloat16)
out = torch.empty(M, 12288, dtype=torch.bfloat16, device="cuda")
for mode in (1, 0):
    _lib.lib.mmdp_set_gemm_pair(mode)
    for _ in range(3):
        _lib.gemm_bf16(a, w, _lib.EPI_PLAIN, out=out)
    torch.cuda.synchronize()
