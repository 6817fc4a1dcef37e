"""Diagnostic: tiny-model forward on the B200 vs (a) the reference's golden logits (CPU) and (b) the oracle forward
executed with torch CUDA kernels (cuBLAS / aten) - i.e. the noise floor between two legitimate implementations."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_golden, tiny_gpu_model
from oracle import llada

g = load_golden("forward_tiny.pt")
model, cfg, sd = tiny_gpu_model(g["meta"])
lg = model(g["ids"], infer=True).logits[0].cpu()
want = g["logits_cols"].float()
got = lg[:, g["cols"]].float()
sd_gpu = {k: v.cuda() for k, v in sd.items()}
lt = llada.forward_logits(g["ids"].cuda(), sd_gpu, cfg)[0].cpu()
gott = lt[:, g["cols"]].float()
def st(a, b):
    d = (a - b).abs()
    return dict(max=float(d.max()), mean=float(d.mean()), p999=float(d.flatten().kthvalue(int(d.numel() * 0.999)).values), scale=float(b.abs().max()), std=float(b.std()))
print("mine_vs_cpu_ref ", json.dumps(st(got, want)))
print("torchgpu_vs_cpu ", json.dumps(st(gott, want)))
print("mine_vs_torchgpu", json.dumps(st(got, gott)))
am = lg.argmax(-1); at = lt.argmax(-1)
print("argmax agree mine/cpu", float((am == g["argmax"]).float().mean()), "torchgpu/cpu", float((at == g["argmax"]).float().mean()), "mine/torchgpu", float((am == at).float().mean()))
margin = (g["top2_vals"][:, 0] - g["top2_vals"][:, 1]).float()
print("ref top1-top2 margin: min", float(margin.min()), "median", float(margin.median()), "  bf16 ulp at top", float(g["top2_vals"][:, 0].float().abs().mean()) * 2 ** -8)
# per-layer drift: hidden after each block (oracle on GPU vs CPU)
