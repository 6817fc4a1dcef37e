"""ncu target: the attention kernel generations on the bench shape (B=1, L=2414, H=32), a few launches each."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmada_parallel_b200 import _lib
torch.manual_seed(0)
B, L, H = 1, 2414, 32
d, M, Lpad = H * 128, B * L, 2416
q = torch.randn(M, d, device="cuda").to(torch.bfloat16)
k = torch.randn(M, d, device="cuda").to(torch.bfloat16)
vt = torch.zeros(B, H, 128, Lpad, dtype=torch.bfloat16, device="cuda")
vt[..., :L] = torch.randn(B, H, 128, L, device="cuda").to(torch.bfloat16)
for ver in [int(v) for v in os.environ.get("MMDP_PROF_VERSIONS", "7,6").split(",")]:
    _lib.check(_lib.lib.mmdp_set_option(b"attn_version", ver))
    for _ in range(3):
        _lib.attention(q, k, vt, B, H, L, 1.0 / math.sqrt(128.0))
    torch.cuda.synchronize()
