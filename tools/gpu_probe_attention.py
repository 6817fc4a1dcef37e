"""Timing decomposition of attention v7 on the bench shape (results of the probe modes are garbage; only times matter):
probe 0 the kernel, 1 softmax threads skip their work (pure MMA chain), 2 no PV MMAs, 3 no QK MMAs."""
import math, os, sys, json
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
flops = 4.0 * B * H * L * L * 128
for rnd in range(2):
    for ver, probe in ((7, 0), (7, 5), (7, 4), (7, 1), (6, 0), (6, 5), (6, 4)):
        _lib.check(_lib.lib.mmdp_set_option(b"attn_version", ver))
        _lib.check(_lib.lib.mmdp_set_option(b"attn_probe", probe))
        for _ in range(3):
            _lib.attention(q, k, vt, B, H, L, 1.0 / math.sqrt(128.0))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _lib.attention(q, k, vt, B, H, L, 1.0 / math.sqrt(128.0))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"version": ver, "probe": probe, "us": round(ms * 1e3, 1), "tflops_equiv": round(flops / ms / 1e9, 1)}), flush=True)
_lib.check(_lib.lib.mmdp_set_option(b"attn_probe", 0))
_lib.check(_lib.lib.mmdp_set_option(b"attn_version", 6))
