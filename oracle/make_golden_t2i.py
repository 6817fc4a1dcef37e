"""Pins oracle.generate.generate_image against the REAL T2I MaskGit decoder of variant A
(MMaDA-Parallel-A/generators/image_generation_generator.py:15-251, imported read-only) and writes
tests/golden/trajectory_t2i_tiny.pt.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_t2i
"""
from __future__ import annotations

import importlib
import os

import torch

from . import generate as G
from . import llada
from . import ref_shim
from .make_golden import OUT, TINY, WEIGHT_SEED, quiet


def layout_t2i(prompt_len=8, grid=4, seed=3):
    g = torch.Generator().manual_seed(seed)
    BOA, BOI, EOI, EOA, MASK, NL = 126354, 126349, 126350, 126355, 126336, 126084
    prompt = torch.randint(0, 126000, (prompt_len,), generator=g).tolist()
    ids = prompt + [BOA, BOI]
    code_start = len(ids)
    for _ in range(grid):
        ids += [MASK] * grid + [NL]
    ids += [EOI, EOA]
    uncon = torch.randint(0, 126000, (1, 3), generator=g)
    return dict(prompt=torch.tensor([ids]), code_start=code_start, seq_len=grid * grid, newline_every=grid, uncon_ids=uncon)


def main():
    assert ref_shim.available(), "reference tree not found"
    torch.set_num_threads(8)
    cfg = llada.make_config(**TINY)
    sd = llada.make_weights(cfg, seed=WEIGHT_SEED)
    with quiet():
        ref = ref_shim.build_ref_model_a(cfg, sd)
    ref_shim.load_a()
    gi = importlib.import_module("generators.image_generation_generator")
    oracle_model = llada.OracleModel(cfg, sd)
    lay = layout_t2i()
    runs = []
    for name, kw, seed in [
        ("greedy_nocfg", dict(timesteps=5, temperature=0.0, cfg_scale=0.0), 1),
        ("temp1_nocfg", dict(timesteps=6, temperature=1.0, cfg_scale=0.0), 2),
        ("temp1_cfg3", dict(timesteps=6, temperature=1.0, cfg_scale=3.0), 9),
        ("early_exit", dict(timesteps=18, temperature=0.7, cfg_scale=1.0), 4),
    ]:
        common = dict(seq_len=lay["seq_len"], newline_every=lay["newline_every"], code_start=lay["code_start"],
                      uncon_ids=lay["uncon_ids"], text_vocab_size=126356, codebook_size=8192, **kw)
        with quiet():
            vr = gi.generate_image(ref, lay["prompt"], generator=torch.Generator().manual_seed(seed), use_cache=False, debug=False, **common)
        # the reference's token cache (use_cache=True): K/V/logits are STORED per block but generate_image never passes
        # `to_compute_mask`, so nothing is ever re-used - pinned here: identical output with the flag on
        with quiet():
            vc = gi.generate_image(ref, lay["prompt"], generator=torch.Generator().manual_seed(seed), use_cache=True, debug=False, **common)
            ref.caching(False)
        assert torch.equal(vr, vc), f"generate_image {name}: use_cache=True changed the reference's output"
        trace = []
        vo = G.generate_image(oracle_model, lay["prompt"], generator=torch.Generator().manual_seed(seed), trace=trace, **common)
        assert torch.equal(vr, vo), f"generate_image {name}: oracle != reference"
        runs.append(dict(name=name, kwargs=kw, seed=seed, vq_ids=vr.clone(), steps_run=len(trace), use_cache_invariant=True))
        print("t2i", name, "ok; steps run:", len(trace), "masked left:", int((vr == 126336).sum()))
    torch.save(dict(meta=dict(tiny=TINY, weight_seed=WEIGHT_SEED), layout=lay, runs=runs), os.path.join(OUT, "trajectory_t2i_tiny.pt"))


if __name__ == "__main__":
    main()
