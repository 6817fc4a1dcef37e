"""ORACLE (test infrastructure, NOT product code): CPU restatement of the two generation loops.

  generate_ti2ti       restates MMaDA-Parallel-A/generators/parallel_generator.py:102-368
  interleave_generate  restates MMaDA-Parallel-M/models/modeling_mmada.py:118-248
  generate_ti2ti_stepwise / decode_text_with_masks  restate MMaDA-Parallel-A/app.py:143-398 / :102-140 (Gradio preview loop)
  mmu_generate         restates MMaDA-Parallel-M/models/modeling_mmada.py:619-691 (semi-autoregressive text generation)
  generate_image       restates MMaDA-Parallel-A/generators/image_generation_generator.py:15-251 (MaskGit T2I; its sampling
                       helpers are those of A/utils/generation_utils.py:28-61, restated next to it)
The per-step arithmetic lives in oracle/sampling.py; this file restates the orchestration (schedules, which forwards
run, how ids are rewritten). The python `.item()` loops of the reference are replaced by tensor indexing with the
same results. Random draws come from sampling.NoiseSource (same calls, same order as the reference).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import sampling as S


def _ti2ti_step(model, ids, step, is_img, k_transfer, pos, noise, text_start, text_end, seq_len, text_steps, temperature,
                text_temperature, cfg_scale, cfg_img, uncon_text, uncon_image, noise_schedule, text_vocab_size, codebook_size,
                stable_sort):
    """One iteration of the step loop (parallel_generator.py:177-344 == app.py:177-305): cond forward, text step, and on
    image steps the uncond forwards + image step. Rewrites `ids` in place; returns the trace record."""
    MASK = S.MASK_TOKEN_A
    cond_logits = model(ids, infer=True, use_cache=False).logits               # :178
    noise.dtype = cond_logits.dtype
    text_ids = ids[0, text_start:text_end]
    if int((text_ids == MASK).sum()) > 0:                                      # :183
        tl = cond_logits[0, text_start:text_end, :]
        un = noise.text_uniform((1,) + tuple(tl.shape))[0].to(tl.device) if text_temperature != 0 else None
        new_ids, x0, conf = S.text_step(tl, text_ids, MASK, k_transfer, temperature=text_temperature, uniform_noise=un)
        ids[0, text_start:text_end] = new_ids
    rec = {"step": step, "ids_after_text": ids[0].clone()}
    if is_img:                                                                 # :220
        tok = ids[0, pos]
        vq = torch.where(tok == MASK, torch.tensor(-1), torch.clamp(tok - text_vocab_size, 0, codebook_size - 1))
        sl = slice(text_vocab_size, text_vocab_size + codebook_size)
        cond_vq = cond_logits[0, pos][:, sl]
        unc_t = unc_i = None
        if (cfg_scale > 0.0 and uncon_text is not None) or (cfg_img > 0.0 and uncon_image is not None):  # :243
            ids_t, ids_i = ids.clone(), ids.clone()
            if uncon_text is not None:
                ids_t[:, : uncon_text.shape[1]] = uncon_text
            if uncon_image is not None:
                ids_i[:, : uncon_image.shape[1]] = uncon_image
            # the reference runs both forwards (:263-264); the uncond_text result is unused when cfg_scale == 0
            if cfg_scale != 0.0:
                unc_t = model(ids_t, infer=True, use_cache=False).logits[0, pos][:, sl]
            unc_i = model(ids_i, infer=True, use_cache=False).logits[0, pos][:, sl]
        else:
            unc_t = torch.zeros_like(cond_vq)                                  # :277-278
            unc_i = torch.zeros_like(cond_vq)
        q = noise.multinomial_q(seq_len, codebook_size).to(cond_vq.device) if temperature != 0 else None
        ratio = 1.0 * (step + 1) / text_steps
        img_temp = temperature * (1.0 - ratio)                                 # :330
        rn = noise.remask_randn((1, seq_len))[0].to(cond_vq.device)            # :333 -> :30-31 (always drawn)
        out = S.image_step("A", cond_vq, unc_t, unc_i, cfg_scale, cfg_img, vq, MASK,
                           S.sched_len(seq_len, step, text_steps, noise_schedule), img_temp, q, rn, codebook_size,
                           stable=stable_sort)
        fin = out["final"]
        ids[0, pos] = torch.where(fin == -1, torch.tensor(MASK), fin + text_vocab_size)   # :339-344
        rec.update(mask_len=out["mask_len"], sampled=out["sampled"].clone(), masking=out["masking"].clone(),
                   ids_after_image=ids[0].clone())
    return rec


@torch.no_grad()
def generate_ti2ti(model, input_ids, text_start, text_end, image_start, seq_len, newline_every, text_steps=100,
                   text_gen_length=256, text_block_length=64, timesteps=100, temperature=1.0, text_temperature=0.7,
                   cfg_scale=0.0, cfg_img=4.0, uncon_text=None, uncon_image=None, tokenizer=None,
                   remasking="low_confidence", noise_schedule=S.cosine_schedule, generator=None,
                   text_vocab_size=126356, codebook_size=8192, trace: Optional[list] = None,
                   final_fill: Optional[torch.Tensor] = None, stable_sort: bool = False):
    """Returns (image_tokens: List[int], text_tokens: List[int] | str). `final_fill` (optional int tensor) supplies
    the values the reference draws with torch.randint on the GLOBAL CPU RNG for still-masked image tokens (:362);
    when None the global RNG is used exactly like the reference."""
    if remasking != "low_confidence":
        raise NotImplementedError(remasking)
    MASK, NL = S.MASK_TOKEN_A, S.NEW_LINE_A
    ids = input_ids.clone()                                                       # :140
    total_image_len = seq_len + seq_len // newline_every
    image_end = image_start + total_image_len
    n_masked = int((ids[0, text_start:text_end] == MASK).sum())
    num_transfer = S.get_num_transfer_tokens_a(n_masked, text_steps)               # :153-154 (batch row 0)
    img_steps = S.image_step_indices(text_steps, timesteps)                       # :157-159
    pos_map = [i for i in range(image_start, image_end) if int(ids[0, i]) != NL]   # :164-167
    assert len(pos_map) == seq_len, f"Expected {seq_len} VQ tokens, got {len(pos_map)}"
    pos = torch.tensor(pos_map, dtype=torch.long)
    noise = S.NoiseSource(generator, dtype=torch.bfloat16)
    for step in range(text_steps):
        rec = _ti2ti_step(model, ids, step, step in img_steps, num_transfer[step], pos, noise, text_start, text_end, seq_len,
                          text_steps, temperature, text_temperature, cfg_scale, cfg_img, uncon_text, uncon_image,
                          noise_schedule, text_vocab_size, codebook_size, stable_sort)
        if trace is not None:
            trace.append(rec)
    text_tokens = [t for t in ids[0, text_start:text_end].tolist() if t != MASK]    # :348-349
    generated_text = tokenizer.decode(text_tokens, skip_special_tokens=True) if tokenizer is not None else text_tokens
    image_tokens, fill_i = [], 0
    for t in ids[0, pos].tolist():                                                 # :353-362
        if t != MASK:
            image_tokens.append(max(0, min(t - text_vocab_size, codebook_size - 1)))
        elif final_fill is not None:
            image_tokens.append(int(final_fill[fill_i]))
            fill_i += 1
        else:
            image_tokens.append(int(torch.randint(0, codebook_size, (1,)).item()))
    return image_tokens, generated_text


class PieceTokenizer:
    """Stand-in tokenizer for pinning / tests (no tokenizer files offline): one printable piece per id; some ids decode to
    a blank, to nothing, or raise, to exercise every branch of decode_text_with_masks (app.py:122-131)."""

    def decode(self, ids, skip_special_tokens=False, clean_up_tokenization_spaces=False):
        t = int(ids[0])
        if t % 11 == 0:
            return " "
        if t % 13 == 0:
            return ""
        if t % 17 == 0:
            raise KeyError(t)
        return f"<{t}>"


def decode_text_with_masks(ids, text_start, text_end, tokenizer, mask_token) -> str:
    """app.py:102-140: text span as a string, runs of masks rendered as blocks (long runs abbreviated)."""
    def run(n):
        return "\u2593" * n if n <= 10 else f"\u2593\u2593\u2593\u2593\u2593[...{n - 5} more]"
    parts, masks = [], 0
    for t in ids[0, text_start:text_end].cpu().tolist():
        if t == mask_token:
            masks += 1
            continue
        if masks > 0:
            parts.append(run(masks))
            masks = 0
        try:
            txt = tokenizer.decode([t], skip_special_tokens=False, clean_up_tokenization_spaces=False)
            if txt.strip() or txt in [" ", "\n", "\t"]:
                parts.append(txt)
        except Exception:                                                          # app.py:131 (bare except)
            parts.append(f"[{t}]")
    if masks > 0:
        parts.append(run(masks))
    return "".join(parts)


def stepwise_image_step_indices(text_steps: int) -> list:
    """app.py:162-164: 30 % of the steps, spread over the whole schedule."""
    return torch.linspace(0, text_steps - 1, int(text_steps * 0.3)).round().int().tolist()


def generate_ti2ti_stepwise(model, input_ids, text_start, text_end, image_start, seq_len, newline_every, text_steps=100,
                            temperature=1.0, text_temperature=0.7, cfg_scale=0.0, cfg_img=4.0, uncon_text=None,
                            uncon_image=None, tokenizer=None, remasking="low_confidence", noise_schedule=S.cosine_schedule,
                            generator=None, text_vocab_size=126356, codebook_size=8192, preview=None,
                            stable_sort: bool = False, trace: Optional[list] = None):
    """Generator restating app.py:143-398. Yields (step, text_display, image, status) at the reference's cadence.
    `preview(sampled_ids[1, N] int64, masking[N] bool | None, masked_idx list | None)` stands for decode_vq_to_image +
    the grey overlay of still-masked cells (app.py:307-335, :367-396); its return value is yielded as the image."""
    if remasking != "low_confidence":
        raise NotImplementedError(remasking)
    MASK, NL = S.MASK_TOKEN_A, S.NEW_LINE_A
    ids = input_ids.clone()
    total_image_len = seq_len + seq_len // newline_every
    n_masked = int((ids[0, text_start:text_end] == MASK).sum())
    num_transfer = S.get_num_transfer_tokens_a(n_masked, text_steps)
    img_steps = stepwise_image_step_indices(text_steps)
    pos_map = [i for i in range(image_start, image_start + total_image_len) if int(ids[0, i]) != NL]
    pos = torch.tensor(pos_map, dtype=torch.long)
    noise = S.NoiseSource(generator, dtype=torch.bfloat16)
    last_image = None
    yield 0, decode_text_with_masks(ids, text_start, text_end, tokenizer, MASK), None, f"Step 0/{text_steps}"
    for step in range(text_steps):
        is_img = step in img_steps
        rec = _ti2ti_step(model, ids, step, is_img, num_transfer[step], pos, noise, text_start, text_end, seq_len, text_steps,
                          temperature, text_temperature, cfg_scale, cfg_img, uncon_text, uncon_image, noise_schedule,
                          text_vocab_size, codebook_size, stable_sort)
        if trace is not None:
            trace.append(rec)
        if is_img and preview is not None:
            last_image = preview(rec["sampled"].unsqueeze(0), rec["masking"], None)   # decode of the PRE-remask sample
        text_display = decode_text_with_masks(ids, text_start, text_end, tokenizer, MASK)
        remaining = int((ids[0, text_start:text_end] == MASK).sum())
        status = f"Step {step + 1}/{text_steps} | Text: {(1 - remaining / (text_end - text_start)) * 100:.1f}%"
        if is_img:
            img_left = int((ids[0, pos] == MASK).sum())
            status += f" | Image: {(1 - img_left / seq_len) * 100:.1f}%"
        if step % 5 == 0 or is_img or step == text_steps - 1:                        # app.py:345
            yield step + 1, text_display, last_image, status
    final_text = decode_text_with_masks(ids, text_start, text_end, tokenizer, MASK)
    if last_image is None and preview is not None:                                   # app.py:352-396 (no image step ran)
        tok = ids[0, pos]
        masked = (tok == MASK).nonzero().flatten().tolist()
        vq = torch.where(tok == MASK, torch.tensor(codebook_size // 2), torch.clamp(tok - text_vocab_size, 0, codebook_size - 1))
        last_image = preview(vq.unsqueeze(0), None, masked)
    yield text_steps, final_text, last_image, "\u2713 Complete"


@torch.no_grad()
def interleave_generate(model, input_ids, uncond_input_ids, text_cfg, image_cfg, text_steps, image_steps, soi_id, eoi_id,
                        bos_id, mask_id, num_vq_tokens, codebook_size, max_seq_length, text_vocab_len,
                        noise_schedule=S.cosine_schedule, generator=None, text_temperature=0.0,
                        image_temperature=1.0, trace: Optional[list] = None, text_noise=None):
    """Restates MMadaModelLM.interleave_generate (modeling_mmada.py:118-248) with the config/tokenizer look-ups
    (reserved_token_mapping, config.model.mmada.*, len(uni_prompting.text_tokenizer)) passed as plain integers.
    `model(ids[B,L]).logits`. Returns (image ids [1, num_vq_tokens] (pre-remask sample), text ids [1, max_seq_length])."""
    if not (text_cfg or image_cfg):
        raise ValueError("text_cfg and image_cfg cannot be both 0")
    # text_temperature != 0: add_gumbel_noise draws fp64 noise [1, max_seq_length, V] from the GLOBAL RNG (modeling_mmada.py:56),
    # replayed here with the same call (seed the global RNG to compare); `text_noise` may inject it instead
    inp, unc = input_ids.unsqueeze(0), uncond_input_ids.unsqueeze(0)
    out_ids = torch.cat([torch.full((1, 1), soi_id), torch.full((1, num_vq_tokens), mask_id), torch.full((1, 1), eoi_id),
                         torch.full((1, 1), bos_id), torch.full((1, max_seq_length - 1), mask_id)], dim=1)  # :142-148
    ids = torch.cat([inp, out_ids], dim=1)
    P = inp.shape[1]
    n_masked = int((ids[:, -max_seq_length:] == mask_id).sum())
    num_transfer = S.get_num_transfer_tokens_m(n_masked, text_steps)
    img_idx = S.image_step_indices(text_steps, image_steps)
    noise = S.NoiseSource(generator)
    sampled_ids = None
    isl = slice(P + 1, P + 1 + num_vq_tokens)
    csl = slice(text_vocab_len, text_vocab_len + codebook_size)
    for i in range(text_steps):
        unc_ids = torch.cat([unc, ids[:, P:]], dim=1)                                # :166-169
        logits = model(torch.cat([ids, unc_ids], dim=0)).logits                     # :172  (B = 2)
        noise.dtype = logits.dtype
        cond_logits, uncond_logits = torch.chunk(logits, 2, dim=0)
        u64 = None
        if text_temperature != 0:
            shape = (1, max_seq_length, logits.shape[-1])
            u64 = (text_noise(i, shape) if text_noise is not None else torch.rand(shape, dtype=torch.float64, device=logits.device))[0]
        new_ids, x0, conf = S.text_step(cond_logits[0, -max_seq_length:], ids[0, -max_seq_length:], mask_id,
                                        num_transfer[i], uncond_logits=uncond_logits[0, -max_seq_length:],
                                        text_cfg=text_cfg, temperature=text_temperature, uniform64=u64)
        ids[0, -max_seq_length:] = new_ids
        rec = {"step": i, "ids_after_text": ids[0].clone()}
        if i in img_idx:                                                             # :211
            tok = ids[0, isl]
            vq = torch.where(tok == mask_id, torch.tensor(mask_id), tok - text_vocab_len)   # :213-214
            q = noise.multinomial_q(num_vq_tokens, codebook_size)
            ratio = 1.0 * (i + 1) / text_steps
            temp = image_temperature * (1.0 - ratio)
            un = noise.remask_uniform((1, num_vq_tokens))[0]
            out = S.image_step("M", cond_logits[0, isl][:, csl], uncond_logits[0, isl][:, csl], None, image_cfg, 0.0, vq,
                               mask_id, S.sched_len(num_vq_tokens, i, text_steps, noise_schedule), temp, q, un,
                               codebook_size)
            sampled_ids = out["sampled"].unsqueeze(0)
            ids[0, isl] = torch.where(out["masking"], torch.tensor(mask_id), out["sampled"] + text_vocab_len)  # :238
            rec.update(mask_len=out["mask_len"], sampled=out["sampled"].clone(), ids_after_image=ids[0].clone())
        if trace is not None:
            trace.append(rec)
    return sampled_ids, ids[:, -max_seq_length:]


@torch.no_grad()
def mmu_generate(model, idx, max_new_tokens=128, steps=128, block_length=128, temperature=0.0, cfg_scale=0.0,
                 remasking="low_confidence", mask_id=126336, attention_mask=None, trace: Optional[list] = None, text_noise=None):
    """Restates MMadaModelLM.mmu_generate (modeling_mmada.py:619-691): LLaDA block-wise un-masking of `max_new_tokens`
    masks appended to the prompt `idx [B, P]`. `model(ids[B', L]).logits`. Per step: forward (CFG: batch [x; x with the
    prompt masked], logits = un + (cfg + 1) * (l - un)), argmax, fp64 softmax confidence, positions after the current block
    excluded, the k most confident masked positions of each row committed. Returns x [B, P + max_new_tokens]."""
    # attention_mask: the reference turns a mask with zeros into `attention_bias` (:626-627) and hands it to self(...) only in
    # the cfg_scale == 0 branch (:663); LLaDAModel.forward of variant M never reads attention_bias (its blocks take
    # `attention_mask`, which stays None) - padding is NOT masked and the argument cannot change the result. Pinned against
    # the real reference with a zero-containing mask in oracle/make_golden_mmu.py.
    # temperature != 0: add_gumbel_noise draws fp64 noise of the FULL logits shape [B, L, V] from the global RNG (:665 -> :56);
    # the same call is issued here (seed the global RNG to compare), or `text_noise(step_index, shape)` injects it.
    if remasking != "low_confidence":
        raise NotImplementedError(remasking)                                         # 'random' uses the global RNG
    B, P = idx.shape
    x = torch.full((B, P + max_new_tokens), mask_id, dtype=torch.long)
    x[:, :P] = idx
    prompt_index = x != mask_id
    assert max_new_tokens % block_length == 0
    num_blocks = max_new_tokens // block_length
    assert steps % num_blocks == 0
    steps = steps // num_blocks
    for blk in range(num_blocks):
        bs, be = P + blk * block_length, P + (blk + 1) * block_length
        counts = (x[:, bs:be] == mask_id).sum(dim=1).tolist()
        num_transfer = [S.get_num_transfer_tokens_m(int(c), steps) for c in counts]
        for i in range(steps):
            if cfg_scale > 0.0:
                un_x = x.clone()
                un_x[prompt_index] = mask_id
                logits, un_logits = torch.chunk(model(torch.cat([x, un_x], dim=0)).logits, 2, dim=0)
            else:
                logits, un_logits = model(x).logits, None
            u64 = None
            if temperature != 0:
                shape = tuple(logits.shape)
                u64 = text_noise(blk * steps + i, shape) if text_noise is not None else torch.rand(shape, dtype=torch.float64, device=logits.device)
            for j in range(B):
                # positions [be, L) carry confidence -inf (:673): only [0, be) can be selected
                uj = u64[j, :be] if u64 is not None else None
                if un_logits is not None:   # un + (cfg + 1) * (l - un)  (:660)
                    new_ids, x0, conf = S.text_step(un_logits[j, :be], x[j, :be], mask_id, num_transfer[j][i],
                                                    uncond_logits=logits[j, :be], text_cfg=cfg_scale + 1, temperature=temperature, uniform64=uj)
                else:
                    new_ids, x0, conf = S.text_step(logits[j, :be], x[j, :be], mask_id, num_transfer[j][i], temperature=temperature, uniform64=uj)
                x[j, :be] = new_ids
            if trace is not None:
                trace.append(x.clone())
    return x


@torch.no_grad()
def t2i_generate(model, input_ids, uncond_input_ids=None, attention_mask=None, uncond_attention_mask=None, temperature=1.0,
                 timesteps=18, guidance_scale=0, noise_schedule=S.cosine_schedule, generator=None, seq_len=1024,
                 mask_token_id=126336, resolution=512, codebook_size=8192, text_vocab_len=126349, trace: Optional[list] = None):
    """Restates MMadaModelLM.t2i_generate (modeling_mmada.py:265-359), MaskGit decoding of the LAST seq_len + 1 positions
    (image tokens followed by one closing token) of `input_ids [B, L]`, which is updated IN PLACE like the reference's tensor.
    The attention masks only build an `attention_bias` the M backbone never reads (see mmu_generate): padding is not masked.
    Quirk kept: `temperature` is multiplied by (1 - ratio) on EVERY step (:341), i.e. it compounds.
    Returns the last step's sampled ids [B, seq_len] (before re-masking)."""
    n, tv, C = seq_len, text_vocab_len, codebook_size
    B = input_ids.shape[0]
    cur = input_ids[:, -(n + 1):-1].clone()
    cur = torch.where(cur == mask_token_id, mask_token_id, cur - tv)                      # :293-294
    if uncond_input_ids is not None:
        uncond_prefix = uncond_input_ids[:, :resolution + 1]                               # :298
    noise = S.NoiseSource(generator)
    sampled_ids = None
    for step in range(timesteps):
        if uncond_input_ids is not None and guidance_scale > 0:                            # :302-316
            uncond_input_ids = torch.cat([uncond_prefix, input_ids[:, resolution + 1:]], dim=1)
            logits = model(torch.cat([input_ids, uncond_input_ids])).logits
            cond_logits, uncond_logits = torch.chunk(logits, 2, dim=0)
            logits = (1 + guidance_scale) * cond_logits - guidance_scale * uncond_logits
        else:
            logits = model(input_ids).logits                                               # :318-319
        logits = logits[:, -(n + 1):-1, tv: tv + C]
        noise.dtype = logits.dtype
        probs = logits.softmax(dim=-1)                                                     # :324
        q = noise.multinomial_q(B * n, C).to(probs.device)                                 # torch.multinomial(probs [B*n, C], 1)  :327
        sampled_ids = S.sample_rows(probs.reshape(-1, C), q).view(B, n)
        unknown_map = cur == mask_token_id
        sampled_ids = torch.where(unknown_map, sampled_ids, cur)                           # :332
        ratio = 1.0 * (step + 1) / timesteps
        mask_ratio = noise_schedule(torch.tensor(ratio))
        selected = torch.gather(probs, -1, sampled_ids.long()[..., None]).squeeze(-1)
        selected = torch.where(unknown_map, selected, torch.finfo(selected.dtype).max)     # :343
        mask_len = (n * mask_ratio).floor().unsqueeze(0)
        mask_len = torch.max(torch.tensor([1]), torch.min(unknown_map.sum(dim=-1, keepdim=True) - 1, mask_len))   # :347-349
        temperature = temperature * (1.0 - ratio)                                          # :352 (compounds)
        un = noise.remask_uniform((B, n)).to(probs.device)
        masking = torch.stack([S.mask_by_random_topk_m(int(mask_len[b]), selected[b], temperature, un[b])[0] for b in range(B)])
        input_ids[:, -(n + 1):-1] = torch.where(masking, mask_token_id, sampled_ids + tv)  # :355-357
        cur = torch.where(masking, mask_token_id, sampled_ids)
        if trace is not None:
            trace.append(dict(step=step, sampled=sampled_ids.clone(), ids=input_ids.clone(), mask_len=mask_len.clone()))
    return sampled_ids


# ---------------------------------------------------------------------------------------------------------------
# A/generators/image_generation_generator.py (T2I MaskGit decoding). No call site in the reference's own scripts; pinned here
# so that the product path can be built against it next.
# ---------------------------------------------------------------------------------------------------------------
def _gumbel_noise_g(t: torch.Tensor, generator=None) -> torch.Tensor:
    """A/utils/generation_utils.py:28-34: -log(-log(u + 1e-20) + 1e-20), u = torch.rand in t's dtype."""
    u = torch.rand_like(t) if generator is None else torch.rand(t.shape, device=t.device, dtype=t.dtype, generator=generator)
    return -torch.log(-torch.log(u + 1e-20) + 1e-20)


def _gumbel_max_sample_g(logits: torch.Tensor, tau: float, generator=None) -> torch.Tensor:
    """generation_utils.py:37-42."""
    if tau == 0.0:
        return logits.argmax(dim=-1)
    return (logits / tau + _gumbel_noise_g(logits, generator)).argmax(dim=-1)


def _mask_by_random_topk_g(mask_len: torch.Tensor, probs: torch.Tensor, temperature: float, generator=None) -> torch.Tensor:
    """generation_utils.py:45-61: True = stay masked; strict `<` against the k-th smallest confidence."""
    confidence = torch.log(probs.clamp_min(1e-20)) + temperature * _gumbel_noise_g(probs, generator)
    sorted_conf = torch.sort(confidence, dim=-1).values
    k = mask_len.long().unsqueeze(1).clamp_(0, probs.size(1) - 1)
    return confidence < torch.gather(sorted_conf, 1, k)


@torch.no_grad()
def generate_image(model, prompt, *, seq_len=1024, newline_every=16, timesteps=18, mask_token_id=126336, newline_id=126084,
                   temperature=1.0, cfg_scale=0.0, uncon_ids=None, code_start=None, codebook_size=8192,
                   noise_schedule=S.cosine_schedule, text_vocab_size=None, generator=None, trace: Optional[list] = None):
    """Restates generate_image with use_cache=False (the token cache is SURVEY 8f rank 4). `model(ids, infer=True).logits`.
    Returns LongTensor [1, seq_len] of full-vocabulary ids (VQ id + text_vocab_size), newlines removed (:236-238)."""
    B, P = prompt.shape
    assert B == 1, "batch>1 not supported – wrap in loop if needed"                    # :57
    x = prompt.clone()
    vq_mask = x == mask_token_id
    unknown_cnt = vq_mask.sum(dim=1, keepdim=True)
    vq_len = unknown_cnt
    if text_vocab_size is None:                                                        # :78-82
        text_vocab_size = model(torch.zeros(1, 1, dtype=torch.long), infer=True).logits.size(-1) - codebook_size
    off = text_vocab_size
    for step in range(timesteps):
        if int(unknown_cnt) == 0:                                                      # :92
            break
        if step < timesteps - 1:                                                       # :99-103
            frac = noise_schedule(torch.tensor([(step + 1) / timesteps]))
            keep_n = (vq_len.float() * frac).floor().clamp_min(1).long()
        else:
            keep_n = torch.zeros_like(unknown_cnt)
        if cfg_scale > 0:                                                              # :121-156
            uncond = torch.cat((uncon_ids, x[:, code_start - 2:]), dim=1)
            uncond_vq_mask = torch.cat((torch.zeros((1, uncon_ids.size(1)), dtype=torch.bool), vq_mask[:, code_start - 2:]), dim=1)
            cond_logits = model(x, infer=True).logits[..., off: off + codebook_size]
            cond_mask_logits = cond_logits[vq_mask].view(B, -1, codebook_size)
            uncond_logits = model(uncond, infer=True).logits[..., off: off + codebook_size]
            uncond_mask_logits = uncond_logits[uncond_vq_mask].view(B, -1, codebook_size)
            logits = (1 + cfg_scale) * cond_mask_logits - cfg_scale * uncond_mask_logits
        else:
            logits = model(x, infer=True).logits[:, vq_mask[0], off: off + codebook_size]   # :162
        sampled = _gumbel_max_sample_g(logits, temperature, generator)                 # :171
        sampled_full = sampled + off
        probs = torch.softmax(logits, dim=-1)
        conf = probs.gather(-1, sampled.unsqueeze(-1)).squeeze(-1)
        flat_idx = vq_mask.nonzero(as_tuple=False)[:, 1]
        x.view(-1)[flat_idx] = sampled_full.view(-1)                                   # :189
        mask_sel = _mask_by_random_topk_g(keep_n.squeeze(1), conf, temperature, generator)   # :204
        x.view(-1)[flat_idx[mask_sel.view(-1)]] = mask_token_id
        vq_mask = x == mask_token_id
        unknown_cnt = vq_mask.sum(dim=1, keepdim=True)
        if trace is not None:
            trace.append(dict(step=step, keep_n=int(keep_n), sampled=sampled.clone(), x=x[0].clone()))
    vq_ids = x[0, code_start:-2]
    return vq_ids[vq_ids != newline_id].view(1, seq_len)
