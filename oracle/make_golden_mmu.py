"""Pins oracle.generate.mmu_generate against the REAL MMadaModelLM.mmu_generate (MMaDA-Parallel-M/models/modeling_mmada.py:
619-691, imported read-only) and writes tests/golden/trajectory_mmu_tiny.pt.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_mmu
"""
from __future__ import annotations

import os

import torch

from . import generate as G
from . import llada
from . import ref_shim
from .make_golden import OUT, TINY, WEIGHT_SEED, quiet


def main():
    assert ref_shim.available(), "reference tree not found"
    torch.set_num_threads(8)
    cfg = llada.make_config(**TINY)
    sd = llada.make_weights(cfg, seed=WEIGHT_SEED)
    mm, _, _ = ref_shim.load_m()
    mcfg = mm.MMadaConfig(**{k: v for k, v in ref_shim.ref_config_a(cfg).to_dict().items()
                             if k not in ("architectures", "model_type", "transformers_version", "mask_token_id")}, mask_token_id=126336)
    mcfg.use_cache = False
    with quiet():
        refm = mm.MMadaModelLM(mcfg, init_params=False).eval().to(torch.bfloat16)
    missing, unexpected = refm.load_state_dict(sd, strict=False)
    assert not unexpected
    oracle_model = llada.OracleModel(cfg, sd)
    g = torch.Generator().manual_seed(21)
    runs = []
    for name, B, P, kw in [
        ("b1_two_blocks", 1, 24, dict(max_new_tokens=16, steps=8, block_length=8, cfg_scale=0.0)),
        ("b2_cfg", 2, 20, dict(max_new_tokens=12, steps=6, block_length=12, cfg_scale=1.5)),
        ("b1_cfg_uneven_k", 1, 17, dict(max_new_tokens=10, steps=4, block_length=10, cfg_scale=0.7)),
        ("b1_all_ones_mask", 1, 16, dict(max_new_tokens=8, steps=4, block_length=4, cfg_scale=0.0, ones_mask=True)),
    ]:
        idx = torch.randint(0, 126000, (B, P), generator=g)
        kw = dict(kw)
        am = torch.ones(B, P + kw["max_new_tokens"], dtype=torch.long) if kw.pop("ones_mask", False) else None
        with quiet():
            xr = refm.mmu_generate(idx=idx, attention_mask=am, **kw)
        trace = []
        xo = G.mmu_generate(oracle_model, idx, attention_mask=am, trace=trace, **kw)
        assert torch.equal(xr, xo), f"mmu_generate {name}: oracle != reference"
        assert int((xo == 126336).sum()) == 0
        runs.append(dict(name=name, idx=idx, kwargs=kw, ones_mask=am is not None, out=xr.clone(), trace=trace))
        print("mmu", name, "ok", tuple(xr.shape))
    torch.save(dict(meta=dict(tiny=TINY, weight_seed=WEIGHT_SEED), runs=runs), os.path.join(OUT, "trajectory_mmu_tiny.pt"))


if __name__ == "__main__":
    main()
