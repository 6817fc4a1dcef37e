"""Pins oracle.llada.CachedOracleModel against the REAL reference's token-cache forward (LLaDAModelLM.forward with
use_cache / to_compute_mask / cat, MMaDA-Parallel-A/model/modeling_llada.py:929-940, :1244-1245, :1406-1413) and
writes tests/golden/token_cache_tiny.pt. No generator of the reference passes a mask; the forward is called directly.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_cache
"""
from __future__ import annotations

import os

import torch

from . import llada
from . import ref_shim
from .make_golden import OUT, TINY, WEIGHT_SEED, quiet


def main():
    assert ref_shim.available(), "reference tree not found"
    torch.set_num_threads(8)
    cfg = llada.make_config(**TINY)
    sd = llada.make_weights(cfg, seed=WEIGHT_SEED)
    with quiet():
        ref = ref_shim.build_ref_model_a(cfg, sd)
    _, _, _, ml = ref_shim.load_a()
    ref.caching(True)
    om = llada.CachedOracleModel(cfg, sd)
    g = torch.Generator().manual_seed(77)
    cases = []
    # batch 1 only: with a mask the reference indexes the rotary table with nonzero()[1] of the WHOLE [B, L] mask (:715), which has
    # B * T' entries for T' query rows - it raises for B > 1
    for name, B, L, cat in [("cond_75", 1, 75, "cond"), ("uncond_40", 1, 40, "uncond")]:
        ids0 = torch.randint(0, 134656, (B, L), generator=g)
        steps = [dict(ids=ids0, mask=None)]
        cur = ids0
        for tq in (9, 17, 1):
            mask = torch.zeros(B, L, dtype=torch.bool)
            for b in range(B):
                mask[b, torch.randperm(L, generator=g)[:tq]] = True
            cur = cur.clone()
            cur[mask] = torch.randint(0, 134656, (int(mask.sum()),), generator=g)      # the recomputed tokens changed
            steps.append(dict(ids=cur, mask=mask))
        recs = []
        for st in steps:
            with torch.no_grad():
                lr = ml.LLaDAModelLM.forward(ref, input_ids=st["ids"], use_cache=True, to_compute_mask=st["mask"], cat=cat).logits
            lo = om(st["ids"], to_compute_mask=st["mask"], cat=cat).logits
            assert torch.equal(lr, lo), f"token cache {name}: oracle != reference"
            cols = torch.randperm(134656, generator=torch.Generator().manual_seed(5))[:96].sort().values
            top = lr.float().topk(2, dim=-1)
            recs.append(dict(ids=st["ids"], mask=st["mask"], logits_cols=lr[:, :, cols].clone(), cols=cols, argmax=lr.float().argmax(-1),
                             top2=top.values.clone()))
        # a partial forward is NOT the dense forward of the new ids (the un-recomputed tokens keep stale keys / values / logits)
        dense = llada.OracleModel(cfg, sd)(steps[-1]["ids"]).logits
        assert not torch.equal(dense, lr)
        cases.append(dict(name=name, cat=cat, steps=recs))
        print("token cache", name, "ok:", len(recs), "forwards")
    torch.save(dict(meta=dict(tiny=TINY, weight_seed=WEIGHT_SEED), cases=cases), os.path.join(OUT, "token_cache_tiny.pt"))


if __name__ == "__main__":
    main()
