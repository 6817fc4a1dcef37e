"""Pins oracle.generate.generate_ti2ti_stepwise / decode_text_with_masks against the REAL reference preview loop
(MMaDA-Parallel-A/app.py:143-398, imported read-only with gradio/diffusers stubbed) and writes
tests/golden/trajectory_stepwise_tiny.pt.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_stepwise

The reference decodes previews with the aMUSEd VQ-VAE (diffusers, not in this image); `decode_vq_to_image` and the module
global `VQVAE` are replaced by recorders, so what is pinned is everything up to the decode call: every yielded
(step, text, status), the ids handed to the decoder at every image step, and the cells the grey overlay marks.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import torch
from PIL import Image

from . import generate as G
from . import llada
from . import ref_shim
from .make_golden import OUT, TINY, WEIGHT_SEED, layout_a, quiet


def main():
    assert ref_shim.available(), "reference tree not found"
    torch.set_num_threads(8)
    cfg = llada.make_config(**TINY)
    sd = llada.make_weights(cfg, seed=WEIGHT_SEED)
    with quiet():
        ref = ref_shim.build_ref_model_a(cfg, sd)
    app = ref_shim.load_app()
    oracle_model = llada.OracleModel(cfg, sd)
    lay = layout_a()
    args = {k: lay[k] for k in ("text_start", "text_end", "image_start", "seq_len", "newline_every", "uncon_text", "uncon_image")}
    grid, scale = 4, 16
    tok = G.PieceTokenizer()
    out = dict(meta=dict(tiny=TINY, weight_seed=WEIGHT_SEED), layout=lay, image_hw=grid * scale, runs=[])
    for name, kw, seed in [
        ("greedy_cfgimg4", dict(text_steps=10, temperature=0.0, text_temperature=0.0, cfg_scale=0.0, cfg_img=4.0), 42),
        ("temp1_both_cfg", dict(text_steps=12, temperature=1.0, text_temperature=0.7, cfg_scale=1.5, cfg_img=4.0), 7),
        ("no_image_step", dict(text_steps=3, temperature=1.0, text_temperature=0.0, cfg_scale=0.0, cfg_img=4.0), 5),
    ]:
        # ---- reference, with the decoder replaced by a recorder
        decoded = []

        def fake_decode(vq_codes, save_path, vae_ckpt, h, w, vqvae):
            decoded.append(vq_codes.clone())
            return Image.new("RGB", (w, h), (10 * len(decoded) % 255, 0, 0))

        app.decode_vq_to_image = fake_decode
        app.VQVAE = SimpleNamespace(config=SimpleNamespace(block_out_channels=[0] * 5))
        ref_y = []
        for step, text, image, status in app.generate_ti2ti_stepwise(
                ref, lay["input_ids"], generator=torch.Generator().manual_seed(seed), tokenizer=tok, vqvae=object(),
                image_height=grid * scale, image_width=grid * scale, **args, **kw):
            overlay = None
            if image is not None:
                px = image.load()  # grey overlay cells differ from the recorder's flat colour
                overlay = [i for i in range(grid * grid)
                           if px[(i % grid) * scale + 1, (i // grid) * scale + 1] != (10 * len(decoded) % 255, 0, 0)]
            ref_y.append((step, text, status, image is not None, overlay))
        # ---- oracle
        o_dec = []

        def preview(sampled, masking, masked_idx):
            o_dec.append(sampled.clone())
            cells = masking.nonzero().flatten().tolist() if masking is not None else list(masked_idx)
            return ("img", len(o_dec), cells)

        trace = []
        or_y = []
        for step, text, image, status in G.generate_ti2ti_stepwise(
                oracle_model, lay["input_ids"], generator=torch.Generator().manual_seed(seed), tokenizer=tok,
                preview=preview, trace=trace, **args, **kw):
            or_y.append((step, text, status, image is not None, None if image is None else image[2]))
        assert len(ref_y) == len(or_y), (name, len(ref_y), len(or_y))
        for a, b in zip(ref_y, or_y):
            assert a[:4] == b[:4], (name, a, b)
        # overlay cells: the reference draws them on the LAST decoded image only at the yield right after an image step
        last = {}
        for a, b in zip(ref_y, or_y):
            if a[4] is not None and a[4] != last.get("ref"):
                assert a[4] == b[4], (name, "overlay", a[0], a[4], b[4])
            last["ref"] = a[4]
        assert len(decoded) == len(o_dec) and all(torch.equal(x, y) for x, y in zip(decoded, o_dec)), name
        out["runs"].append(dict(name=name, kwargs=kw, seed=seed, yields=[(a[0], a[1], a[2], a[3]) for a in ref_y],
                                decoded=[d.clone() for d in decoded],
                                overlays=[b[4] for b in or_y],
                                final_ids=trace[-1].get("ids_after_image", trace[-1]["ids_after_text"]).clone()))
        print("stepwise", name, "ok:", len(ref_y), "yields,", len(decoded), "decodes")
    torch.save(out, os.path.join(OUT, "trajectory_stepwise_tiny.pt"))


if __name__ == "__main__":
    main()
