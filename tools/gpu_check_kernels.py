"""Kernel-level bring-up check on a real B200 (run through gpurun). Each case runs in its own subprocess with a
timeout so that a hung kernel cannot take the whole call down. Writes gpurun_out/kernels_check.json + log.

    python tools/gpu_check_kernels.py            # all cases
    python tools/gpu_check_kernels.py --case gemm_plain_small
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bf16(x):
    import torch
    return x.to(torch.bfloat16)


def ref_linear(a, w):
    """nn.Linear rounding contract: fp32 accumulate -> bf16."""
    return _bf16(a.float() @ w.float().t())


def stats(name, got, want, atol=0.0, rtol=0.0, extra=None):
    import torch
    g, w = got.float(), want.float()
    diff = (g - w).abs()
    denom = w.abs().clamp_min(1e-6)
    out = {
        "case": name,
        "max_abs": float(diff.max()),
        "max_rel": float((diff / denom).max()),
        "mismatch_frac": float((g != w).float().mean()),
        "nan": bool(torch.isnan(g).any()),
        "want_absmax": float(w.abs().max()),
    }
    out["ok"] = (not out["nan"]) and bool(((diff <= atol + rtol * w.abs())).all())
    if extra:
        out.update(extra)
    return out


def time_it(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / iters


# ------------------------------------------------------------------------------------------------------------
def case_gemm(M, N, K, epi="plain", seed=0, timing=False):
    import torch
    from mmada_parallel_b200 import _lib
    torch.manual_seed(seed)
    dev = "cuda"
    a = _bf16(torch.randn(M, K, device=dev) * 0.5)
    w = _bf16(torch.randn(N, K, device=dev) * 0.05)
    res = []
    if epi == "plain":
        got = _lib.gemm_bf16(a, w, _lib.EPI_PLAIN)
        want = ref_linear(a, w)
    elif epi == "resid":
        r = _bf16(torch.randn(M, N, device=dev))
        got = _lib.gemm_bf16(a, w, _lib.EPI_RESID, resid=r)
        want = _bf16(ref_linear(a, w).float() + r.float())
    elif epi == "resid_inplace":
        r = _bf16(torch.randn(M, N, device=dev))
        want = _bf16(ref_linear(a, w).float() + r.float())
        got = _lib.gemm_bf16(a, w, _lib.EPI_RESID, resid=r, out=r)
    elif epi == "swiglu":
        ff = N // 2
        w1, w3 = w[:ff], w[ff:]
        # pack: 128-row blocks interleaved
        wp = torch.empty_like(w)
        wp.view(ff // 128, 2, 128, K)[:, 0] = w1.view(ff // 128, 128, K)
        wp.view(ff // 128, 2, 128, K)[:, 1] = w3.view(ff // 128, 128, K)
        got = _lib.gemm_bf16(a, wp, _lib.EPI_SWIGLU)
        g = ref_linear(a, w1)
        u = ref_linear(a, w3)
        s = _bf16(torch.nn.functional.silu(g.float()))
        want = _bf16(s.float() * u.float())
    else:
        raise ValueError(epi)
    torch.cuda.synchronize()
    # bf16 output: allow 1 bf16 ulp (accumulation-order differences move a value across a rounding boundary)
    out = stats(f"gemm_{epi}_{M}x{N}x{K}", got, want, atol=2e-3, rtol=1.0 / 128)
    if timing:
        if epi == "plain":
            ms = time_it(lambda: _lib.gemm_bf16(a, w, _lib.EPI_PLAIN, out=got))
            ms_t = time_it(lambda: torch.matmul(a, w.t()))
            out["ms"] = ms
            out["tflops"] = 2.0 * M * N * K / ms / 1e9
            out["torch_ms"] = ms_t
            out["torch_tflops"] = 2.0 * M * N * K / ms_t / 1e9
    return out


def time_resid(M, N, K):
    import torch
    from mmada_parallel_b200 import _lib
    a = _bf16(torch.randn(M, K, device="cuda") * 0.5)
    w = _bf16(torch.randn(N, K, device="cuda") * 0.05)
    r = _bf16(torch.randn(M, N, device="cuda"))
    out = torch.empty_like(r)
    ms = time_it(lambda: _lib.gemm_bf16(a, w, _lib.EPI_RESID, resid=r, out=out))
    ms_t = time_it(lambda: torch.addmm(r, a, w.t()))
    return {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "torch_addmm_ms": ms_t, "torch_tflops": 2.0 * M * N * K / ms_t / 1e9}


def ref_rope(t, cos, sin):
    """t: [M, H, 128] bf16; cos/sin [M, 64] fp32 -> reference apply_rotary_pos_emb in fp32."""
    import torch
    tf = t.float()
    x1, x2 = tf[..., :64], tf[..., 64:]
    c = cos[:, None, :]
    s = sin[:, None, :]
    o1 = x1 * c + (-x2) * s
    o2 = x2 * c + x1 * s
    return _bf16(torch.cat([o1, o2], dim=-1))


def case_qkv_attn(B, L, H, seed=0, timing=False):
    import torch
    from mmada_parallel_b200 import _lib
    from mmada_parallel_b200.model import rope_tables
    torch.manual_seed(seed)
    dev = "cuda"
    d = H * 128
    M = B * L
    a = _bf16(torch.randn(M, d, device=dev))
    wqkv = _bf16(torch.randn(3 * d, d, device=dev) * (1.0 / math.sqrt(d)))
    cos, sin = rope_tables(128, 500000.0, L)
    cos, sin = cos.to(dev), sin.to(dev)
    q, k, vt = _lib.qkv_rope(a, wqkv, H, L, cos, sin)
    torch.cuda.synchronize()
    qkv = ref_linear(a, wqkv)
    pos = torch.arange(M, device=dev) % L
    q_ref = ref_rope(qkv[:, :d].view(M, H, 128), cos[pos], sin[pos]).view(M, d)
    k_ref = ref_rope(qkv[:, d:2 * d].view(M, H, 128), cos[pos], sin[pos]).view(M, d)
    v_ref = qkv[:, 2 * d:]
    res = [stats(f"qkv_q_B{B}L{L}H{H}", q, q_ref, atol=4e-3, rtol=1.0 / 64),
           stats(f"qkv_k_B{B}L{L}H{H}", k, k_ref, atol=4e-3, rtol=1.0 / 64)]
    v_got = vt[..., :L].permute(0, 3, 1, 2).reshape(M, d)  # [B,H,128,L] -> [B,L,H,128]
    res.append(stats(f"qkv_vt_B{B}L{L}H{H}", v_got, v_ref, atol=2e-3, rtol=1.0 / 128,
                     extra={"pad_zero": bool((vt[..., L:] == 0).all())}))
    # attention on the library's own q/k/vt against torch SDPA in fp32
    scale = 1.0 / math.sqrt(128.0)
    o = _lib.attention(q, k, vt, B, H, L, scale)
    torch.cuda.synchronize()
    qh = q.view(B, L, H, 128).transpose(1, 2).float()
    kh = k.view(B, L, H, 128).transpose(1, 2).float()
    vh = v_got.reshape(B, L, H, 128).transpose(1, 2).float()
    o_ref = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1) @ vh
    o_ref = o_ref.transpose(1, 2).reshape(M, d)
    res.append(stats(f"attention_B{B}L{L}H{H}", o, o_ref, atol=2e-2, rtol=2e-2))
    if timing:
        ms = time_it(lambda: _lib.attention(q, k, vt, B, H, L, scale))
        res[-1]["ms"] = ms
        res[-1]["tflops"] = 4.0 * B * H * L * L * 128 / ms / 1e9
        ms_q = time_it(lambda: _lib.qkv_rope(a, wqkv, H, L, cos, sin))
        res[0]["ms"] = ms_q
        res[0]["tflops"] = 2.0 * M * 3 * d * d / ms_q / 1e9
    return res


def case_rmsnorm(M, d, seed=0):
    import torch
    from mmada_parallel_b200 import _lib
    torch.manual_seed(seed)
    x = _bf16(torch.randn(M, d, device="cuda") * 3)
    w = _bf16(1 + 0.1 * torch.randn(d, device="cuda"))
    y = _lib.rmsnorm(x, w, 1e-5)
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    want = w * _bf16(xf * torch.rsqrt(var + 1e-5))
    rows = torch.tensor([3, 1, M - 1, 0], dtype=torch.int32, device="cuda")
    y2 = _lib.rmsnorm(x, w, 1e-5, rows=rows)
    return [stats(f"rmsnorm_{M}x{d}", y, want, atol=1e-6, rtol=1.0 / 128),
            stats(f"rmsnorm_rows_{M}x{d}", y2, want[rows.long()], atol=1e-6, rtol=1.0 / 128)]


def case_embed_lfq():
    import torch
    from mmada_parallel_b200 import _lib
    torch.manual_seed(0)
    wte = _bf16(torch.randn(1000, 256, device="cuda"))
    ids = torch.randint(0, 1000, (77,), device="cuda")
    x = _lib.embed(ids, wte)
    out = [stats("embed", x, wte[ids])]
    vq = torch.randint(0, 8192, (2, 1024), device="cuda")
    zq = _lib.lfq_decode(vq, 13)
    binary = (vq.unsqueeze(-1) >> torch.arange(12, -1, -1, device="cuda")) & 1
    want = (binary.float() * 2 - 1).permute(0, 2, 1)
    out.append(stats("lfq_decode", zq, want))
    return out


def case_pair(M, N, K, epi="plain", timing=True):
    """CTA-pair (cta_group::2) kernel vs the 1-CTA kernel: same K order -> must be bit-identical."""
    import torch
    from mmada_parallel_b200 import _lib
    torch.manual_seed(M + N)
    dev = "cuda"
    a = _bf16(torch.randn(M, K, device=dev) * 0.5)
    w = _bf16(torch.randn(N, K, device=dev) * 0.05)
    r = _bf16(torch.randn(M, N, device=dev))
    code = {"plain": _lib.EPI_PLAIN, "resid": _lib.EPI_RESID, "swiglu": _lib.EPI_SWIGLU}[epi]
    kw = dict(resid=r) if epi == "resid" else {}
    _lib.lib.mmdp_set_gemm_pair(0)
    ref = _lib.gemm_bf16(a, w, code, **kw)
    torch.cuda.synchronize()
    _lib.lib.mmdp_set_gemm_pair(1)
    got = _lib.gemm_bf16(a, w, code, **kw)
    torch.cuda.synchronize()
    out = stats(f"pair_{epi}_{M}x{N}x{K}", got, ref)
    out["bit_identical"] = bool(torch.equal(got, ref))
    out["ok"] = out["bit_identical"]
    if timing:
        ms_p = time_it(lambda: _lib.gemm_bf16(a, w, code, out=got, **kw))
        _lib.lib.mmdp_set_gemm_pair(0)
        ms_s = time_it(lambda: _lib.gemm_bf16(a, w, code, out=got, **kw))
        out.update(pair_ms=ms_p, single_ms=ms_s, pair_tflops=2.0 * M * N * K / ms_p / 1e9, single_tflops=2.0 * M * N * K / ms_s / 1e9)
    return out


def case_pair_qkv(B, L, H):
    import torch
    from mmada_parallel_b200 import _lib
    from mmada_parallel_b200.model import rope_tables
    torch.manual_seed(5)
    d, M = H * 128, B * L
    a = _bf16(torch.randn(M, d, device="cuda"))
    wqkv = _bf16(torch.randn(3 * d, d, device="cuda") / math.sqrt(d))
    cos, sin = (t.cuda() for t in rope_tables(128, 500000.0, L))
    _lib.lib.mmdp_set_gemm_pair(0)
    q0, k0, v0 = _lib.qkv_rope(a, wqkv, H, L, cos, sin)
    torch.cuda.synchronize()
    _lib.lib.mmdp_set_gemm_pair(1)
    q1, k1, v1 = _lib.qkv_rope(a, wqkv, H, L, cos, sin)
    torch.cuda.synchronize()
    ok = bool(torch.equal(q0, q1) and torch.equal(k0, k1) and torch.equal(v0, v1))
    out = {"case": f"pair_qkv_B{B}L{L}H{H}", "bit_identical": ok, "ok": ok}
    ms_p = time_it(lambda: _lib.qkv_rope(a, wqkv, H, L, cos, sin))
    _lib.lib.mmdp_set_gemm_pair(0)
    ms_s = time_it(lambda: _lib.qkv_rope(a, wqkv, H, L, cos, sin))
    out.update(pair_ms=ms_p, single_ms=ms_s, pair_tflops=2.0 * M * 3 * d * d / ms_p / 1e9, single_tflops=2.0 * M * 3 * d * d / ms_s / 1e9)
    return out


CASES = {
    "splitk_resid_attnout": lambda: case_gemm(2414, 4096, 4096, "resid", timing=False) | time_resid(2414, 4096, 4096),
    "splitk_resid_ffout": lambda: case_gemm(2414, 4096, 12288, "resid", timing=False) | time_resid(2414, 4096, 12288),
    "splitk_resid_small": lambda: case_gemm(300, 512, 768, "resid"),
    "splitk_resid_b2": lambda: case_gemm(4828, 4096, 4096, "resid"),
    "pair_small": lambda: case_pair(512, 512, 256, timing=False),
    "pair_ragged": lambda: case_pair(777, 1000, 520, timing=False),
    "pair_resid": lambda: case_pair(2414, 4096, 4096, "resid"),
    "pair_ffout": lambda: case_pair(2414, 4096, 12288, "resid"),
    "pair_swiglu": lambda: case_pair(2414, 24576, 4096, "swiglu"),
    "pair_plain_qkvshape": lambda: case_pair(2414, 12288, 4096),
    "pair_qkv": lambda: case_pair_qkv(1, 2414, 32),
    "gemm_plain_tile": lambda: case_gemm(128, 256, 64),
    "gemm_plain_k": lambda: case_gemm(128, 256, 512),
    "gemm_plain_multi": lambda: case_gemm(512, 1024, 1024),
    "gemm_plain_ragged": lambda: case_gemm(333, 264, 200),
    "gemm_resid": lambda: case_gemm(300, 512, 768, "resid"),
    "gemm_resid_inplace": lambda: case_gemm(300, 512, 768, "resid_inplace"),
    "gemm_swiglu": lambda: case_gemm(300, 1024, 512, "swiglu"),
    "gemm_persistent": lambda: case_gemm(2414, 4096, 4096, timing=True),
    "gemm_qkv_shape": lambda: case_gemm(2414, 12288, 4096, timing=True),
    "gemm_ffout_shape": lambda: case_gemm(2414, 4096, 12288, timing=True),
    "gemm_head_text": lambda: case_gemm(256, 134656, 4096, timing=True),
    "resid_attnout_b3": lambda: {"case": "resid_7242x4096x4096", "ok": True, **time_resid(7242, 4096, 4096)},
    "resid_ffout_b3": lambda: {"case": "resid_7242x4096x12288", "ok": True, **time_resid(7242, 4096, 12288)},
    "gemm_swiglu_b1": lambda: case_gemm(2414, 24576, 4096, timing=True),
    "gemm_qkv_b3": lambda: case_gemm(7242, 12288, 4096, timing=True),
    "gemm_swiglu_b3": lambda: case_gemm(7242, 24576, 4096, timing=True),
    "gemm_odd_m": lambda: case_gemm(2414 + 128, 1024, 512, timing=False),
    "qkv_attn_small": lambda: case_qkv_attn(1, 128, 2),
    "qkv_attn_ragged": lambda: case_qkv_attn(2, 200, 2),
    "qkv_attn_multi": lambda: case_qkv_attn(1, 640, 4),
    "qkv_attn_real": lambda: case_qkv_attn(1, 2414, 32, timing=True),
    "rmsnorm": lambda: case_rmsnorm(300, 4096),
    "rmsnorm_small": lambda: case_rmsnorm(17, 256),
    "embed_lfq": case_embed_lfq,
}


def run_case(name):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    r = CASES[name]()
    if isinstance(r, dict):
        r = [r]
    print("RESULT " + json.dumps(r))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--timeout", type=int, default=150)
    ap.add_argument("--only", default=None, help="comma-separated prefixes")
    args = ap.parse_args()
    if args.case:
        run_case(args.case)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results, log = [], []
    names = [n for n in CASES if not args.only or any(n.startswith(p) for p in args.only.split(","))]
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True, text=True,
                               timeout=args.timeout, cwd=ROOT)
            rc, out, err = p.returncode, p.stdout, p.stderr
        except subprocess.TimeoutExpired as e:
            rc, out, err = -999, (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
        dt = time.time() - t0
        got = [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
        entry = {"name": name, "rc": rc, "sec": round(dt, 1), "results": got[0] if got else None}
        if rc != 0:
            entry["stderr_tail"] = err[-1500:]
            entry["stdout_tail"] = out[-1500:]
        results.append(entry)
        ok = rc == 0 and got and all(r.get("ok") for r in got[0])
        print(f"[{'OK ' if ok else 'BAD'}] {name} rc={rc} {dt:.1f}s " + (json.dumps(got[0]) if got else err[-400:]), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "kernels_check.json"), "w") as f:
            json.dump(results, f, indent=1)
    bad = [e["name"] for e in results if e["rc"] != 0 or not e["results"] or not all(r.get("ok") for r in e["results"])]
    print("SUMMARY bad=" + json.dumps(bad))


if __name__ == "__main__":
    main()
