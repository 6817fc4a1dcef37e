"""Isolated timing of the body GEMMs of the bench workload under tuning options, variants ALTERNATED (clock drift hits all
alike), next to torch.matmul (cuBLAS) on the same shapes. Run through gpurun; writes gpurun_out/probe_gemm.json.

    python tools/gpu_probe_gemm.py                 # M = 2414 (variant A) and 4682 (variant M, CFG batch 2)
"""
from __future__ import annotations

import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "pair": dict(gemm_pair=1, gemm_nsplit_tail=0, gemm_mtail=0),
    "pair_nsplit": dict(gemm_pair=1, gemm_nsplit_tail=1, gemm_mtail=0),
    "pair_all": dict(gemm_pair=1, gemm_nsplit_tail=1, gemm_mtail=1),   # + half-M units (SwiGLU GEMM)
    "single": dict(gemm_pair=0, gemm_nsplit_tail=0, gemm_mtail=0),
}


def main():
    from mmada_parallel_b200 import _lib
    from mmada_parallel_b200.model import rope_tables
    iters, rounds = 20, 3
    out = []
    for M in (2414, 4682):
        d, ff = 4096, 12288
        torch.manual_seed(M)
        a = (torch.randn(M, d, device="cuda") * 0.5).to(torch.bfloat16)
        a_ff = (torch.randn(M, ff, device="cuda") * 0.5).to(torch.bfloat16)
        r = torch.randn(M, d, device="cuda").to(torch.bfloat16)
        w_qkv = (torch.randn(3 * d, d, device="cuda") * 0.02).to(torch.bfloat16)
        w_o = (torch.randn(d, d, device="cuda") * 0.02).to(torch.bfloat16)
        w_gu = (torch.randn(2 * ff, d, device="cuda") * 0.02).to(torch.bfloat16)
        w_dn = (torch.randn(d, ff, device="cuda") * 0.02).to(torch.bfloat16)
        cos, sin = (t.cuda() for t in rope_tables(128, 500000.0, M))
        o_d = torch.empty(M, d, dtype=torch.bfloat16, device="cuda")
        o_ff = torch.empty(M, ff, dtype=torch.bfloat16, device="cuda")
        cases = {
            "qkv_rope": (lambda: _lib.qkv_rope(a, w_qkv, 32, M, cos, sin), 2.0 * M * 3 * d * d, lambda: a @ w_qkv.t()),
            "attn_out": (lambda: _lib.gemm_bf16(a, w_o, _lib.EPI_RESID, resid=r, out=o_d), 2.0 * M * d * d, lambda: a @ w_o.t()),
            "gate_up_swiglu": (lambda: _lib.gemm_bf16(a, w_gu, _lib.EPI_SWIGLU, out=o_ff), 2.0 * M * 2 * ff * d, lambda: a @ w_gu.t()),
            "ff_out": (lambda: _lib.gemm_bf16(a_ff, w_dn, _lib.EPI_RESID, resid=r, out=o_d), 2.0 * M * d * ff, lambda: a_ff @ w_dn.t()),
        }

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        for name, (fn, flops, cublas) in cases.items():
            acc = {v: [] for v in list(VARIANTS) + ["cublas"]}
            for _ in range(rounds):
                for v, opts in VARIANTS.items():
                    for k, val in opts.items():
                        _lib.check(_lib.lib.mmdp_set_option(k.encode(), int(val)))
                    acc[v].append(timed(fn))
                acc["cublas"].append(timed(cublas))
            rec = {"M": M, "case": name}
            for v, ms in acc.items():
                med = statistics.median(ms)
                rec[v + "_us"] = round(med * 1e3, 2)
                rec[v + "_tflops"] = round(flops / med / 1e9, 1)
            out.append(rec)
            print(json.dumps(rec), flush=True)
    for k, val in dict(gemm_pair=1, gemm_nsplit_tail=1, gemm_mtail=1).items():
        _lib.check(_lib.lib.mmdp_set_option(k.encode(), int(val)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_gemm.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
