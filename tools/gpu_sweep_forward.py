"""Times ONE transformer forward of the bench workload (8B shapes, B=1, L=2414, restricted LM head) under different tuning
options (mmdp_set_option) in a single process, and checks that the logits of every variant stay within bf16 rounding of
the first one. Run through gpurun; writes gpurun_out/sweep_forward.json.

    python tools/gpu_sweep_forward.py                       # default variant list
    python tools/gpu_sweep_forward.py --variants base all   # a subset
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BASE = dict(pdl=0, gemm_splitk=1, gemm_l2pf=0, gemm_l2pf_mod=4, attn_split_tail=0, gemm_pair=0, gemm_nsplit_tail=0, gemm_mtail=0)
VARIANTS = {
    "r1": {},                                                         # round-1 configuration
    "cur": dict(pdl=1, gemm_splitk=2, attn_split_tail=1),             # 1-CTA kernel with this round's options
    "cur_nopdl": dict(pdl=0, gemm_splitk=2, attn_split_tail=1),
    "cur_sk1": dict(pdl=1, gemm_splitk=1, attn_split_tail=1),
    "cur_sk0": dict(pdl=1, gemm_splitk=0, attn_split_tail=1),
    "cur_noattn": dict(pdl=1, gemm_splitk=2, attn_split_tail=0),
    "cur_pair": dict(pdl=1, gemm_splitk=2, attn_split_tail=1, gemm_pair=1),
    "cur_pair_nsplit": dict(pdl=1, gemm_splitk=2, attn_split_tail=1, gemm_pair=1, gemm_nsplit_tail=1),
    "cur_all": dict(pdl=1, gemm_splitk=2, attn_split_tail=1, gemm_pair=1, gemm_nsplit_tail=1, gemm_mtail=1),   # current defaults
    "r1_pair": dict(gemm_pair=1),
    "l2pf8": dict(gemm_l2pf=8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", nargs="*", default=list(VARIANTS))
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--tiny", action="store_true")
    args = ap.parse_args()
    import bench
    from mmada_parallel_b200 import _lib
    cfg = bench.MODEL_TINY if args.tiny else bench.MODEL_8B
    model = bench.build_model(cfg, "cuda:0", seed=1000)
    lay = bench.synthetic_layout(seed=0)
    ids = lay["input_ids"].cuda()
    L = ids.shape[1]
    text_rows = torch.arange(lay["text_start"], lay["text_end"], dtype=torch.int32, device="cuda")
    pos = torch.tensor([i for i in range(lay["image_start"], lay["image_start"] + 1024 + 32) if int(ids[0, i]) != bench.NL],
                       dtype=torch.int32, device="cuda")
    out_a = torch.empty((text_rows.numel(), cfg["vocab_size"]), dtype=torch.bfloat16, device="cuda")
    out_b = torch.empty((pos.numel(), bench.CODEBOOK), dtype=torch.bfloat16, device="cuda")

    def fwd():
        model.forward_rows(ids, rows_a=text_rows, out_a=out_a, rows_b=pos, col0_b=bench.TEXT_VOCAB, ncols_b=bench.CODEBOOK, out_b=out_b)

    import statistics
    ref = None
    acc = {name: dict(ms=[], gemm=[], attn=[], row=[]) for name in args.variants}
    checks = {}
    for rnd in range(args.rounds):  # alternate the variants: clock / temperature drift hits all of them alike
        for name in args.variants:
            opts = dict(BASE)
            opts.update(VARIANTS[name])
            for k, v in opts.items():
                _lib.check(_lib.lib.mmdp_set_option(k.encode(), int(v)))
            fwd()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fwd()
            e1.record()
            torch.cuda.synchronize()
            acc[name]["ms"].append(e0.elapsed_time(e1) / args.iters)
            _lib.lib.mmdp_prof_enable(1)
            fwd()
            prof = _lib.prof_summary()
            _lib.lib.mmdp_prof_enable(0)
            acc[name]["gemm"].append(prof["gemm"][0])
            acc[name]["attn"].append(prof["attention"][0])
            acc[name]["row"].append(prof["row"][0])
            if rnd == 0:
                a, b = out_a.float().clone(), out_b.float().clone()
                rec = {"nan": bool(torch.isnan(a).any() or torch.isnan(b).any())}
                if ref is None:
                    ref = (a, b)
                else:
                    rec["max_abs_diff_text"] = float((a - ref[0]).abs().max())
                    rec["max_abs_diff_img"] = float((b - ref[1]).abs().max())
                    rec["mean_abs_diff_text"] = float((a - ref[0]).abs().mean())
                    rec["frac_diff_text"] = float((a != ref[0]).float().mean())
                    rec["logit_absmax"] = float(ref[0].abs().max())
                fwd()
                torch.cuda.synchronize()
                rec["repeatable"] = bool(torch.equal(out_a.float(), a) and torch.equal(out_b.float(), b))
                checks[name] = rec
    results = []
    flops = prof["gemm"][1]
    for name in args.variants:
        m = acc[name]
        rec = {"variant": name, "opts": {**BASE, **VARIANTS[name]}, "ms_per_forward_median": statistics.median(m["ms"]), "ms_per_forward_all": m["ms"],
               "gemm_ms_median": statistics.median(m["gemm"]), "attn_ms_median": statistics.median(m["attn"]), "row_ms_median": statistics.median(m["row"]),
               "gemm_tflops": flops / statistics.median(m["gemm"]) / 1e9, **checks[name]}
        results.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_forward.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
