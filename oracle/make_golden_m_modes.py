"""Pins the remaining variant-M generation modes of oracle.generate against the REAL reference (MMaDA-Parallel-M/models/
modeling_mmada.py, imported read-only) and writes tests/golden/trajectory_m_modes_tiny.pt:

  * t2i_generate (:265-359): B = 1 without guidance, B = 2 with guidance and attention masks that CONTAIN ZEROS (the masks only
    build an attention_bias the M backbone never reads - equality with the all-ones run is asserted on the reference itself);
  * mmu_generate (:619-691) with a zero-containing attention_mask, and with temperature > 0 (fp64 Gumbel from the global RNG,
    re-seeded before both sides);
  * interleave_generate (:118-248) with text_temperature > 0 (same global-RNG Gumbel).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_m_modes
"""
from __future__ import annotations

import os
import types

import torch

from . import generate as G
from . import llada
from . import ref_shim
from .make_golden import OUT, TINY, WEIGHT_SEED, quiet

TV, MASK = 126349, 126336


def main():
    assert ref_shim.available(), "reference tree not found"
    torch.set_num_threads(8)
    cfg = llada.make_config(**TINY)
    sd = llada.make_weights(cfg, seed=WEIGHT_SEED)
    mm, _, _ = ref_shim.load_m()
    mcfg = mm.MMadaConfig(**{k: v for k, v in ref_shim.ref_config_a(cfg).to_dict().items()
                             if k not in ("architectures", "model_type", "transformers_version", "mask_token_id")}, mask_token_id=MASK)
    mcfg.use_cache = False
    with quiet():
        refm = mm.MMadaModelLM(mcfg, init_params=False).eval().to(torch.bfloat16)
    _, unexpected = refm.load_state_dict(sd, strict=False)
    assert not unexpected
    om = llada.OracleModel(cfg, sd)

    class Tok:
        bos_token_id = 126080

        def __len__(self):
            return TV

    up = types.SimpleNamespace(text_tokenizer=Tok())
    g = torch.Generator().manual_seed(31)
    out = dict(meta=dict(tiny=TINY, weight_seed=WEIGHT_SEED, text_vocab_len=TV), t2i=[], mmu=[], interleave=[])

    # ---------------- t2i_generate
    for name, B, n, res, kw, zeros in [("b1_noguidance", 1, 16, 6, dict(timesteps=5, temperature=1.0, guidance_scale=0), False),
                                       ("b2_guidance_zero_masks", 2, 16, 8, dict(timesteps=6, temperature=1.0, guidance_scale=2.5), True),
                                       ("b1_guidance_hot", 1, 25, 5, dict(timesteps=4, temperature=3.0, guidance_scale=1.0), False)]:
        P = res + 1 + 7
        inp = torch.cat([torch.randint(0, 126000, (B, P), generator=g), torch.full((B, n), MASK), torch.full((B, 1), 126086)], dim=1)
        unc = inp.clone()
        unc[:, :res + 1] = torch.randint(0, 126000, (B, res + 1), generator=g)
        am = torch.ones_like(inp)
        if zeros:
            am[0, :3] = 0
            am[1, 2:5] = 0
        common = dict(seq_len=n, resolution=res, mask_token_id=MASK, codebook_size=8192, **kw)
        with quiet():
            a = inp.clone()
            vr = refm.t2i_generate(input_ids=a, uncond_input_ids=unc.clone(), attention_mask=am, uncond_attention_mask=am,
                                   generator=torch.Generator().manual_seed(5), uni_prompting=up, **common)
            if zeros:  # the masks are dead: all-ones masks give the same ids
                a1 = inp.clone()
                v1 = refm.t2i_generate(input_ids=a1, uncond_input_ids=unc.clone(), attention_mask=torch.ones_like(am),
                                       uncond_attention_mask=torch.ones_like(am), generator=torch.Generator().manual_seed(5),
                                       uni_prompting=up, **common)
                assert torch.equal(v1, vr) and torch.equal(a1, a), "attention masks changed the reference's t2i_generate"
        b = inp.clone()
        trace = []
        vo = G.t2i_generate(om, b, unc.clone(), attention_mask=am, uncond_attention_mask=am, generator=torch.Generator().manual_seed(5),
                            text_vocab_len=TV, trace=trace, **common)
        assert torch.equal(vr, vo) and torch.equal(a, b), f"t2i_generate {name}: oracle != reference"
        out["t2i"].append(dict(name=name, input_ids=inp, uncond_input_ids=unc, attention_mask=am, kwargs=common, seed=5,
                               sampled=vr.clone(), final_input_ids=a.clone()))
        print("t2i", name, "ok", tuple(vr.shape), "masks left:", int((a == MASK).sum()))

    # ---------------- mmu_generate: padded mask (dead) and temperature > 0 (global RNG)
    for name, B, P, kw, zeros, gseed in [("b2_zero_mask", 2, 20, dict(max_new_tokens=8, steps=4, block_length=8, cfg_scale=0.0), True, None),
                                         ("b1_temp", 1, 18, dict(max_new_tokens=8, steps=4, block_length=4, cfg_scale=0.0, temperature=0.9), False, 77),
                                         ("b2_temp_cfg", 2, 12, dict(max_new_tokens=6, steps=3, block_length=6, cfg_scale=1.2, temperature=0.5), False, 78)]:
        idx = torch.randint(0, 126000, (B, P), generator=g)
        am = None
        if zeros:
            am = torch.ones(B, P + kw["max_new_tokens"], dtype=torch.long)
            am[0, :4] = 0
        if gseed is not None:
            torch.manual_seed(gseed)
        with quiet():
            xr = refm.mmu_generate(idx=idx, attention_mask=am, **kw)
            if zeros:
                assert torch.equal(xr, refm.mmu_generate(idx=idx, attention_mask=None, **kw)), "padding mask changed the reference's mmu_generate"
        if gseed is not None:
            torch.manual_seed(gseed)
        xo = G.mmu_generate(om, idx, attention_mask=am, **kw)
        assert torch.equal(xr, xo), f"mmu_generate {name}: oracle != reference"
        out["mmu"].append(dict(name=name, idx=idx, attention_mask=am, kwargs=kw, global_seed=gseed, out=xr.clone()))
        print("mmu", name, "ok")

    # ---------------- interleave_generate with text_temperature > 0
    conf = types.SimpleNamespace(model=types.SimpleNamespace(mmada=types.SimpleNamespace(num_vq_tokens=16, codebook_size=8192)),
                                 dataset=types.SimpleNamespace(preprocessing=types.SimpleNamespace(max_seq_length=12)))
    inp = torch.cat([torch.tensor([126340, 126085]), torch.randint(TV, TV + 8192, (16,), generator=g), torch.tensor([126086]),
                     torch.randint(0, 126000, (6,), generator=g)])
    unc = inp.clone()
    unc[-6:] = torch.randint(0, 126000, (6,), generator=g)
    kw = dict(text_cfg=1.5, image_cfg=3.0, text_steps=8, image_steps=4, text_temperature=0.8, image_temperature=1.0)
    torch.manual_seed(91)
    with quiet():
        ir, tr = refm.interleave_generate(input_ids=inp, uncond_input_ids=unc, reserved_token_mapping={"<|soi|>": 126085, "<|eoi|>": 126086},
                                          generator=torch.Generator().manual_seed(8), config=conf, uni_prompting=up, **kw)
    torch.manual_seed(91)
    io, to = G.interleave_generate(om, inp, unc, soi_id=126085, eoi_id=126086, bos_id=126080, mask_id=MASK, num_vq_tokens=16,
                                   codebook_size=8192, max_seq_length=12, text_vocab_len=TV, generator=torch.Generator().manual_seed(8), **kw)
    assert torch.equal(ir, io) and torch.equal(tr, to), "interleave_generate with text_temperature: oracle != reference"
    out["interleave"].append(dict(name="text_temp", input_ids=inp, uncond_input_ids=unc, kwargs=kw, seed=8, global_seed=91,
                                  image_ids=ir.clone(), text_ids=tr.clone()))
    print("interleave text_temperature ok")
    torch.save(out, os.path.join(OUT, "trajectory_m_modes_tiny.pt"))


if __name__ == "__main__":
    main()
