"""Bring-up check of the attention kernel generations on a real B200 (run through gpurun): every case compares the selected
versions (mmdp_set_option attn_version) with the fp32 softmax(QK^T)V reference under the relative bound of
tests/test_gpu_kernels.py::assert_attention_close, checks bitwise repeatability, and times them alternately.
Writes gpurun_out/check_attention.json.

    python tools/gpu_check_attention.py                 # versions 6 and 7
    python tools/gpu_check_attention.py --versions 7 --cases small
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SMALL = [(1, 128, 2, 0.0), (2, 200, 2, 0.0), (1, 640, 4, 0.0), (3, 77, 2, 0.0), (1, 256, 1, 0.0), (1, 257, 3, 0.0), (2, 333, 2, 0.0),
         (1, 1000, 2, 60.0), (1, 129, 1, 200.0), (1, 40, 1, 0.0)]
LARGE = [(1, 2414, 4, 0.0), (1, 2414, 32, 0.0), (2, 2341, 32, 0.0), (1, 2414, 16, 0.0), (1, 4096, 8, 0.0)]


def bf(x):
    return x.to(torch.bfloat16)


def make(B, L, H, grow, seed):
    torch.manual_seed(seed)
    d, M = H * 128, B * L
    Lpad = (L + 7) // 8 * 8
    q = bf(torch.randn(M, d, device="cuda"))
    ramp = 1.0 + grow * (torch.arange(M, device="cuda") % L).float()[:, None] / L
    k = bf(torch.randn(M, d, device="cuda") * ramp)
    v = bf(torch.randn(M, d, device="cuda"))
    vt = torch.zeros(B, H, 128, Lpad, dtype=torch.bfloat16, device="cuda")
    vt[..., :L] = v.view(B, L, H, 128).permute(0, 2, 3, 1)
    return q, k, v, vt


def reference(q, k, v, B, L, H, scale):
    outs = []
    for h0 in range(0, H, 4):  # 4 heads at a time: [B, 4, L, L] fp32 stays small
        hs = slice(h0, min(H, h0 + 4))
        qh = q.view(B, L, H, 128)[:, :, hs].transpose(1, 2).float()
        kh = k.view(B, L, H, 128)[:, :, hs].transpose(1, 2).float()
        vh = v.view(B, L, H, 128)[:, :, hs].transpose(1, 2).float()
        outs.append((torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1) @ vh).transpose(1, 2))
    return torch.cat(outs, dim=2).reshape(B * L, H * 128)


def err_stats(o, o_ref):
    g, w = o.float(), o_ref.float()
    rms = w.pow(2).mean(-1, keepdim=True).sqrt()
    rel = (g - w).abs() / torch.maximum(w.abs(), rms)
    return {"nan": bool(torch.isnan(g).any()), "max_rel_ulp": float(rel.max() * 256), "mean_rel_ulp_of_rms": float(((g - w).abs() / rms).mean() * 256)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--versions", nargs="*", type=int, default=[6, 7])
    ap.add_argument("--cases", default="all")
    ap.add_argument("--repeat", type=int, default=20)
    args = ap.parse_args()
    from mmada_parallel_b200 import _lib
    scale = 1.0 / math.sqrt(128.0)
    cases = SMALL if args.cases == "small" else LARGE if args.cases == "large" else SMALL + LARGE
    out = []
    ok_all = True
    for (B, L, H, grow) in cases:
        q, k, v, vt = make(B, L, H, grow, 1000 * B + L + H)
        o_ref = reference(q, k, v, B, L, H, scale)
        rec = {"B": B, "L": L, "H": H, "grow": grow}
        outs = {}
        for ver in args.versions:
            _lib.check(_lib.lib.mmdp_set_option(b"attn_version", ver))
            for split in (1, 0):
                _lib.check(_lib.lib.mmdp_set_option(b"attn_split_tail", split))
                o = _lib.attention(q, k, vt, B, H, L, scale)
                torch.cuda.synchronize()
                st = err_stats(o, o_ref)
                same = all(torch.equal(_lib.attention(q, k, vt, B, H, L, scale), o) for _ in range(args.repeat if L * H > 20000 else 3))
                st["repeatable"] = bool(same)
                st["ok"] = (not st["nan"]) and st["max_rel_ulp"] <= 4.0 and st["mean_rel_ulp_of_rms"] < 0.5 and same
                ok_all &= st["ok"]
                rec[f"v{ver}_split{split}"] = st
                outs[(ver, split)] = o.float()
            _lib.check(_lib.lib.mmdp_set_option(b"attn_split_tail", 1))
        if len(args.versions) == 2:
            a, b_ = outs[(args.versions[0], 1)], outs[(args.versions[1], 1)]
            rec["max_abs_diff_between_versions"] = float((a - b_).abs().max())
        if L * H * B >= 2414 * 4:
            flops = 4.0 * B * H * L * L * 128
            acc = {ver: [] for ver in args.versions}
            for _ in range(3):
                for ver in args.versions:
                    _lib.check(_lib.lib.mmdp_set_option(b"attn_version", ver))
                    for _ in range(3):
                        _lib.attention(q, k, vt, B, H, L, scale)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        _lib.attention(q, k, vt, B, H, L, scale)
                    e1.record()
                    torch.cuda.synchronize()
                    acc[ver].append(e0.elapsed_time(e1) / 10)
            for ver in args.versions:
                ms = statistics.median(acc[ver])
                rec[f"v{ver}_us"] = round(ms * 1e3, 1)
                rec[f"v{ver}_tflops"] = round(flops / ms / 1e9, 1)
        out.append(rec)
        print(json.dumps(rec), flush=True)
    _lib.check(_lib.lib.mmdp_set_option(b"attn_version", 6))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "check_attention.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("ALL OK" if ok_all else "FAILURES", flush=True)
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
