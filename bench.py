#!/usr/bin/env python
"""Benchmark of the parallel-denoising hot path (BASELINE.json metric):
denoised tokens/sec (text+image) per 512x512 @ 64-step sample, variant A 8B, cfg_img=4 (BASELINE configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                            # the reference algorithm on the host CPU (oracle port)

One "step" = one full sample = one generate_ti2ti call: 128 denoising iterations, 192 transformer forwards
(128 conditional + 64 unconditional-image), 128 text steps, 64 image steps -> 1280 denoised tokens.
Under torchrun (N > 1) every rank denoises its own independent prompt (replicas, no data-path collective): weak scaling.
Prints ONE JSON line on rank 0. Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks. Weights (16.2 GB) are re-read from HBM every forward, far beyond the 126 MB L2, so no L2 flush is needed.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOKENS_PER_SAMPLE = 1024 + 256
MASK, NL, BOA, BOI, EOI, EOA = 126336, 126084, 126354, 126349, 126350, 126355
TEXT_VOCAB, CODEBOOK = 126356, 8192

MODEL_8B = dict(d_model=4096, n_heads=32, n_layers=32, mlp_hidden_size=12288, vocab_size=134656, max_sequence_length=2432)
MODEL_TINY = dict(d_model=256, n_heads=2, n_layers=2, mlp_hidden_size=512, vocab_size=134656, max_sequence_length=2432)
GEN = dict(text_steps=128, timesteps=64, text_gen_length=256, text_block_length=32, temperature=1.0, text_temperature=0.0,
           cfg_scale=0.0, cfg_img=4.0)  # README.md:101-117 of the reference


def synthetic_layout(seed: int, prompt_len: int = 40, grid: int = 32, text_len: int = 256):
    """SURVEY.md 8d synthetic input A: L = P + 2374 = 2414 at P = 40 (structure of A/inference.py:129-156)."""
    g = torch.Generator().manual_seed(seed)
    prompt = torch.randint(0, 126000, (prompt_len,), generator=g).tolist()
    img_in = torch.randint(TEXT_VOCAB, TEXT_VOCAB + CODEBOOK, (grid * grid,), generator=g).tolist()
    img = [BOI]
    for r in range(grid):
        img += img_in[r * grid:(r + 1) * grid] + [NL]
    img += [EOI]
    con = prompt[:-1] + img + prompt[-1:]
    pred = [BOA, BOI]
    for _ in range(grid):
        pred += [MASK] * grid + [NL]
    pred += [EOI] + [MASK] * text_len + [EOA]
    ids = con + pred
    image_start = len(con) + 2
    text_start = image_start + grid * (grid + 1) + 1
    unc_prompt = torch.randint(0, 126000, (3,), generator=g).tolist()
    return dict(input_ids=torch.tensor([ids]), text_start=text_start, text_end=text_start + text_len,
                image_start=image_start, seq_len=grid * grid, newline_every=grid,
                uncon_text=torch.tensor([unc_prompt[:-1] + img + unc_prompt[-1:]]), uncon_image=torch.tensor([prompt]))


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.p, self.path = gpu_index, None, f"/tmp/mmdp_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        loaded = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, bounded sample, extrapolated by exact counts
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(model_cfg: dict, reps: int = 1):
    """Times the path's CPU restatement (oracle/, the port of the reference; /root/reference is not on the GPU box) at
    the full BASELINE shapes on a bounded sample and extrapolates to one full sample with the exact operation counts:
      192 forwards x n_layers block-forwards  +  128 text-row heads  +  128 image-col heads (64 cond + 64 uncond)
      + 128 text steps + 64 image steps."""
    from oracle import llada, sampling as S
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = llada.make_config(**model_cfg)
    d, ff, V, L = cfg.d_model, cfg.mlp_hidden_size, cfg.vocab_size, 2414
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std).to(torch.bfloat16)
    p = "b."
    w = {p + "q_proj.weight": rnd(d, d), p + "k_proj.weight": rnd(d, d), p + "v_proj.weight": rnd(d, d),
         p + "attn_out.weight": rnd(d, d), p + "ff_proj.weight": rnd(ff, d), p + "up_proj.weight": rnd(ff, d),
         p + "ff_out.weight": rnd(d, ff), p + "attn_norm.weight": torch.ones(d, dtype=torch.bfloat16),
         p + "ff_norm.weight": torch.ones(d, dtype=torch.bfloat16)}
    head = rnd(V, d)
    x = rnd(1, L, d, std=1.0)
    pos_sin, pos_cos = llada.rotary_tables(d // cfg.n_heads, cfg.rope_theta, L)
    with torch.no_grad():
        llada.block_forward(x, w, p, cfg, pos_sin, pos_cos)  # warm-up (thread pool, oneDNN primitives)
        t0 = time.perf_counter()
        for _ in range(reps):
            y = llada.block_forward(x, w, p, cfg, pos_sin, pos_cos)
        t_block = (time.perf_counter() - t0) / reps
        xt = y[0, 2157:2413]
        t0 = time.perf_counter()
        tl = torch.nn.functional.linear(xt, head)
        t_head_text = time.perf_counter() - t0
        xi = y[0, 1100:1100 + 1024]
        t0 = time.perf_counter()
        il = torch.nn.functional.linear(xi, head[TEXT_VOCAB:TEXT_VOCAB + CODEBOOK])
        t_head_img = time.perf_counter() - t0
        ids = torch.full((256,), MASK)
        t0 = time.perf_counter()
        S.text_step(tl, ids, MASK, 2)
        t_text = time.perf_counter() - t0
        vq = torch.full((1024,), -1)
        q = torch.empty(1024, CODEBOOK, dtype=torch.bfloat16).exponential_(1, generator=g)
        rn = torch.randn(1024, generator=g).to(torch.bfloat16)
        t0 = time.perf_counter()
        S.image_step("A", il, None, il.flip(0), 0.0, 4.0, vq, MASK, 600, 0.5, q, rn, CODEBOOK)
        t_img = time.perf_counter() - t0
    n_layers = cfg.n_layers
    per_sample = 192 * n_layers * t_block + 128 * t_head_text + 128 * t_head_img + 128 * t_text + 64 * t_img
    return dict(tokens_per_s=TOKENS_PER_SAMPLE / per_sample, sec_per_sample=per_sample, cores=threads,
                t_block=t_block, t_head_text=t_head_text, t_head_img=t_head_img, t_text_step=t_text, t_image_step=t_img,
                sample=f"{reps} of {192 * n_layers} block forwards (d={d}, ff={ff}, L={L}) + text/image heads + 1 text step + 1 image step, "
                       f"extrapolated with the exact per-sample counts")


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    cfg = MODEL_TINY if args.tiny else MODEL_8B
    vals = []
    for _ in range(args.warmup):
        pass  # the CPU sample does its own warm-up forward; extra warm-up samples would only burn minutes
    r = None
    for _ in range(max(1, min(args.steps, 3))):
        r = cpu_reference_sample(cfg)
        vals.append(r["tokens_per_s"])
    v = statistics.median(vals)
    out = {"impl": "reference", "metric": "denoised_tokens_per_sec", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * TOKENS_PER_SAMPLE / v, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": workload_config(args, 1),
           "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
           "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "detail": {k: r[k] for k in ("t_block", "t_head_text", "t_head_img", "t_text_step", "t_image_step", "sec_per_sample")}}
    print(json.dumps(out), flush=True)


def workload_config(args, n_gpus):
    return {"workload": "MMaDA-Parallel-A 8B, 1 prompt per GPU, 512x512 (1024 VQ tokens) + 256 text tokens, timesteps=64, "
                        "text_steps=128, cfg_img=4.0, cfg_scale=0, temperature=1.0, text_temperature=0 (BASELINE configs[1]"
                        + ("; TINY MODEL - plumbing check only, not a valid number" if args.tiny else "") + ")",
            "seq_len": 2414, "forwards_per_sample": 192,
            "parallelism": (f"tensor-parallel x{n_gpus} (one prompt; heads/ff/vocab split, fp32 all-reduce over NCCL)" if getattr(args, "tp", False) and n_gpus > 1
                            else f"replicas x{n_gpus} (independent prompts, no collective)"),
            "weights": "synthetic normal(0, 0.02) bf16, seeded", "l2": "16.2 GB of weights streamed per forward >> 126 MB L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def model_namespace(model_cfg: dict):
    from types import SimpleNamespace
    return SimpleNamespace(**model_cfg, n_kv_heads=None, embedding_size=model_cfg["vocab_size"], rope_theta=500000.0,
                           rms_norm_eps=1e-5, rope=True, rope_full_precision=True, include_bias=False, weight_tying=False)


def synthetic_tensors(model_cfg: dict, device: str, seed: int):
    """Yields (HF name, bf16 tensor on `device`) of a seeded random-init model: normal(0, 0.02) matrices, unit norms."""
    g = torch.Generator(device=device).manual_seed(seed)
    d, ff, V = model_cfg["d_model"], model_cfg["mlp_hidden_size"], model_cfg["vocab_size"]

    def mk(*shape, ones=False):
        if ones:
            return torch.ones(shape, dtype=torch.bfloat16, device=device)
        return torch.empty(shape, dtype=torch.bfloat16, device=device).normal_(0.0, 0.02, generator=g)

    yield "model.transformer.wte.weight", mk(V, d)
    yield "model.transformer.ff_out.weight", mk(V, d)
    yield "model.transformer.ln_f.weight", mk(d, ones=True)
    for i in range(model_cfg["n_layers"]):
        p = f"model.transformer.blocks.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "attn_out"):
            yield p + n + ".weight", mk(d, d)
        yield p + "ff_proj.weight", mk(ff, d)
        yield p + "up_proj.weight", mk(ff, d)
        yield p + "ff_out.weight", mk(d, ff)
        yield p + "attn_norm.weight", mk(d, ones=True)
        yield p + "ff_norm.weight", mk(d, ones=True)


def build_model(model_cfg: dict, device: str, seed: int):
    from mmada_parallel_b200.model import LLaDAForMultiModalGeneration
    m = LLaDAForMultiModalGeneration(model_namespace(model_cfg), max_seq_len=model_cfg["max_sequence_length"], max_batch=1, device=device)
    for name, t in synthetic_tensors(model_cfg, device, seed):
        assert m.set_weight(name, t)
    m.load_state_dict({}, strict=True)
    torch.cuda.synchronize()
    return m


def build_tp_model(model_cfg: dict, device: str, seed: int, rank: int, world: int):
    """Tensor-parallel model (BASELINE config 4): every rank materialises the same seeded tensors and keeps its shard."""
    from mmada_parallel_b200.tensor_parallel import TensorParallelLLaDA
    sd = dict(synthetic_tensors(model_cfg, device, seed))
    m = TensorParallelLLaDA(model_namespace(model_cfg), sd, rank, world, max_seq_len=model_cfg["max_sequence_length"], device=device,
                            text_vocab_size=TEXT_VOCAB, codebook_size=CODEBOOK)
    del sd
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    return m


def algorithmic_flops_per_sample(c: dict, L: int = 2414):
    d, ff, V, nl = c["d_model"], c["mlp_hidden_size"], c["vocab_size"], c["n_layers"]
    body = nl * (2 * L * (4 * d * d + 3 * d * ff) + 4 * L * L * d)
    head_text, head_img = 2 * 256 * d * V, 2 * 1024 * d * CODEBOOK
    return 192 * body + 128 * head_text + 128 * head_img  # minimal-equivalent work


def run_variant_m(args, rank, local_rank, world):
    """Extra measurement (not the contract line): variant M, one prompt per GPU, L = 2341, B = 2 (cond + uncond) on every
    one of the 128 steps, 64 image steps (SURVEY.md 8d synthetic input M). Public API call with host inputs = e2e."""
    from types import SimpleNamespace
    import torch.distributed as dist
    from mmada_parallel_b200.mmada import MMadaModelLM
    from mmada_parallel_b200.parallel import max_over_ranks
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(device))
    model_cfg = MODEL_TINY if args.tiny else MODEL_8B
    ns = model_namespace(model_cfg)
    ns.mask_token_id = MASK
    m = MMadaModelLM(ns, max_seq_len=model_cfg["max_sequence_length"], max_batch=2, device=device)
    for name, t in synthetic_tensors(model_cfg, device, 1000):
        assert m.set_weight(name, t)
    m.load_state_dict({}, strict=True)
    g = torch.Generator().manual_seed(rank)
    tvoc, soi, eoi, bos = 126349, 126085, 126086, 126080
    inp = torch.cat([torch.tensor([126340, soi]), torch.randint(tvoc, tvoc + CODEBOOK, (1024,), generator=g), torch.tensor([eoi]),
                     torch.randint(0, 126000, (32,), generator=g)])
    unc = inp.clone()
    unc[-32:] = torch.randint(0, 126000, (32,), generator=g)
    conf = SimpleNamespace(model=SimpleNamespace(mmada=SimpleNamespace(num_vq_tokens=1024, codebook_size=CODEBOOK)),
                           dataset=SimpleNamespace(preprocessing=SimpleNamespace(max_seq_length=256)))

    class Tok:
        bos_token_id = bos

        def __len__(self):
            return tvoc

    kw = dict(input_ids=inp, uncond_input_ids=unc, text_cfg=2.5, image_cfg=4.0, text_steps=128, image_steps=64,
              reserved_token_mapping={"<|soi|>": soi, "<|eoi|>": eoi}, config=conf, uni_prompting=SimpleNamespace(text_tokenizer=Tok()))
    rng = torch.Generator(device=device).manual_seed(42 + rank)
    with torch.no_grad():
        for _ in range(args.warmup):
            m.interleave_generate(generator=rng, **kw)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            img, txt = m.interleave_generate(generator=rng, **kw)
            txt.cpu()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), device=device)
    if rank == 0:
        v = world * args.steps * TOKENS_PER_SAMPLE / (ms / 1e3)
        print(json.dumps({"metric": "denoised_tokens_per_sec", "value": v, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": "MMaDA-Parallel-M 8B (extra line, not the contract metric): 1 prompt per GPU, L=2341, CFG batch 2 on "
                                                 "each of 128 steps, 64 image steps, text_cfg=2.5, image_cfg=4.0", "parallelism": f"replicas x{world}"},
                          "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": int(inp.numel() * 16), "d2h_bytes_per_step": 256 * 8}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiny", action="store_true", help="2-layer d=256 model: plumbing check only (INVALID as a benchmark number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="a", choices=["a", "m"], help="a = BASELINE configs[1] (the contract metric); m = extra line for "
                    "variant M (interleave_generate, CFG batch 2 every step, BASELINE configs[4] per GPU)")
    ap.add_argument("--tp", action="store_true", help="N > 1: ONE sample tensor-parallel over the N GPUs (strong scaling, NCCL "
                    "all-reduce) instead of N independent replicas")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.variant == "m":
        run_variant_m(args, rank, local_rank, world)
        return

    import torch.distributed as dist
    from mmada_parallel_b200 import _lib
    from mmada_parallel_b200.generators.parallel_generator import DenoiseState, denoise_loop, generate_ti2ti
    from mmada_parallel_b200.parallel import max_over_ranks
    from mmada_parallel_b200.schedule import cosine_schedule

    device = f"cuda:{local_rank}"
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(device))
    model_cfg = MODEL_TINY if args.tiny else MODEL_8B
    tp_mode = args.tp and world > 1
    if tp_mode:
        model = build_tp_model(model_cfg, device, 1000, rank, world)
        lay = synthetic_layout(seed=0)  # ONE prompt, all ranks work on it
    else:
        model = build_model(model_cfg, device, seed=1000)
        lay = synthetic_layout(seed=rank)  # every rank denoises its own prompt
    host_ids = lay["input_ids"].pin_memory()
    pos_args = {k: lay[k] for k in ("text_start", "text_end", "image_start", "seq_len", "newline_every")}
    gen = GEN
    loop_kw = dict(text_steps=gen["text_steps"], timesteps=gen["timesteps"], temperature=gen["temperature"],
                   text_temperature=gen["text_temperature"], cfg_scale=gen["cfg_scale"], cfg_img=gen["cfg_img"],
                   noise_schedule=cosine_schedule, text_vocab_size=TEXT_VOCAB, codebook_size=CODEBOOK)

    def new_state():
        return DenoiseState(model, lay["input_ids"], uncon_text=lay["uncon_text"], uncon_image=lay["uncon_image"],
                            cfg_scale=gen["cfg_scale"], cfg_img=gen["cfg_img"], codebook_size=CODEBOOK, **pos_args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rng = torch.Generator(device=device).manual_seed(42 if tp_mode else 42 + rank)  # TP ranks must draw identical noise
    with torch.no_grad():
        for _ in range(args.warmup):
            denoise_loop(new_state(), generator=rng, **loop_kw)
        # ---- timed region 1: inputs resident in HBM ("value")
        states = [new_state() for _ in range(args.steps)]
        clocks = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            clocks.start()
        _lib.lib.mmdp_launch_count(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for st in states:
            denoise_loop(st, generator=rng, **loop_kw)
        ev1.record()
        barrier()
        launches = int(_lib.lib.mmdp_launch_count(0))
        clock_info = clocks.stop() if rank == 0 else None
        ms_value = max_over_ranks(ev0.elapsed_time(ev1), device=device)

        # ---- timed region 2: through the public API with HOST buffers ("e2e"): H2D of the inputs + D2H of the result inside
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with quiet():
            for _ in range(args.steps):
                img, txt = generate_ti2ti(model, host_ids, uncon_text=lay["uncon_text"], uncon_image=lay["uncon_image"], generator=rng,
                                          text_gen_length=gen["text_gen_length"], text_block_length=gen["text_block_length"],
                                          **pos_args, **{k: v for k, v in loop_kw.items()})
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), device=device)
        h2d = states[0].bytes_h2d()
        d2h = host_ids.numel() * 8

        # ---- roofline pass (rank 0): one sample with every launch bracketed by CUDA events
        prof = None
        if rank == 0 or tp_mode:  # tensor-parallel: every rank must take part (the forward contains collectives)
            _lib.lib.mmdp_prof_enable(1)
            denoise_loop(new_state(), generator=rng, **loop_kw)
            prof = _lib.prof_summary()
            _lib.lib.mmdp_prof_enable(0)
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_samples = args.steps if tp_mode else world * args.steps
    total_tokens = n_samples * TOKENS_PER_SAMPLE
    value = total_tokens / (ms_value / 1e3)
    e2e_value = total_tokens / (ms_e2e / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback ~1.4 PFLOP/s sustained (B200_PROFILING.md)"
    gemm_ms, gemm_flops, gemm_n = prof["gemm"]
    att_ms, att_flops, att_n = prof["attention"]
    row_ms, row_bytes, row_n = prof["row"]
    smp_ms, smp_bytes, smp_n = prof["sampling"]
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("gemm_dram_bytes_per_launch")
    except Exception:
        pass
    flops_sample = algorithmic_flops_per_sample(model_cfg)
    out = {
        "metric": "denoised_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True,
        "scaling": "strong" if tp_mode else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clock_info,
        "roofline": {"kernel": "gemm_bf16_kernel (tcgen05, all epilogues)", "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                     "peak_source": peak_src, "launches": gemm_n, "avg_launch_ms": gemm_ms / max(1, gemm_n),
                     "algorithmic_flops_per_launch": gemm_flops / max(1, gemm_n)},
        "kernel_breakdown_one_sample_ms": {"gemm": gemm_ms, "attention": att_ms, "row_kernels": row_ms, "sampling": smp_ms,
                                           "attention_tflops": att_flops / (att_ms / 1e3) / 1e12 if att_ms else None,
                                           "row_GBps": row_bytes / (row_ms / 1e3) / 1e9 if row_ms else None,
                                           "sampling_GBps": smp_bytes / (smp_ms / 1e3) / 1e9 if smp_ms else None},
        "whole_step_tflops_minimal_work": flops_sample * n_samples / (ms_value / 1e3) / 1e12,
        "whole_step_frac_of_peak": flops_sample * n_samples / (ms_value / 1e3) / 1e12 / (peak_tf * world),  # per-GPU fraction
    }
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_sample(model_cfg)
        out["cpu_baseline"] = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                               "sample": r["sample"], "sec_per_sample": r["sec_per_sample"]}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
