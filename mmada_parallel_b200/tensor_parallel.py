"""Tensor-parallel forward for the single-sample case (BASELINE config 4): one sample's transformer forward split over the
GPUs of one node. The reference has no tensor parallelism; this follows SURVEY.md 8e:

  * attention: heads split (q/k/v_proj column-parallel, attn_out row-parallel); MLP: ff split (ff_proj/up_proj
    column-parallel, ff_out row-parallel); LM head: vocabulary rows split (and the VQ-codebook window split separately);
  * the two row-parallel GEMMs per layer produce fp32 PARTIAL sums (MMDP_EPI_F32) that are all-reduced in fp32 over
    NCCL/NVLink and only then rounded to bf16 and added to the residual (mmdp_resid_add_f32), i.e. the same rounding
    points as the single-GPU epilogue `bf16(bf16(acc) + x)` - a bf16 all-reduce would round twice;
  * every rank then holds the full residual stream, gathers the logits slices it needs (all_gather) and runs the SAME
    sampling kernels on the same noise (identical generator seeds), so the id sequence stays in sync without broadcasts.

The layer loop lives here (Python calling the C-ABI ops of libmmdp.so); the collective is torch.distributed (NCCL).
Fusing the all-reduce into the GEMM epilogue over NVLink peer memory is the planned next step.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import EPI_F32, EPI_PLAIN, EPI_SWIGLU, check, lib, ptr, stream_ptr
from .model import ModelOutput, rope_tables


def shard_state_dict(sd: Dict[str, torch.Tensor], n_layers: int, n_heads: int, rank: int, tp: int, vq_col0: int,
                     vq_cols: int) -> Dict[str, torch.Tensor]:
    """Slices a full HF state dict (names of SURVEY.md 8b) into rank `rank`'s tensor-parallel shard. Pure tensor slicing
    (device agnostic) so it is unit-tested on the CPU."""
    g = lambda n: sd["model.transformer." + n] if ("model.transformer." + n) in sd else sd[n]
    out = {"wte": g("wte.weight"), "ln_f": g("ln_f.weight")}
    head = g("ff_out.weight")
    V, d = head.shape
    if n_heads % tp or V % tp or vq_cols % tp:
        raise ValueError(f"tp={tp} must divide n_heads={n_heads}, vocab rows={V} and the codebook window={vq_cols}")
    da = (n_heads // tp) * 128
    out["head"] = head[rank * (V // tp):(rank + 1) * (V // tp)]
    c = vq_cols // tp
    out["head_vq"] = head[vq_col0 + rank * c: vq_col0 + (rank + 1) * c]
    for i in range(n_layers):
        p = f"blocks.{i}."
        sl = slice(rank * da, (rank + 1) * da)
        out[p + "wqkv"] = torch.cat([g(p + "q_proj.weight")[sl], g(p + "k_proj.weight")[sl], g(p + "v_proj.weight")[sl]], dim=0)
        out[p + "wo"] = g(p + "attn_out.weight")[:, sl]
        ffp, up = g(p + "ff_proj.weight"), g(p + "up_proj.weight")
        ff = ffp.shape[0]
        if (ff // tp) % 128:
            raise ValueError("mlp_hidden / tp must be a multiple of 128 (SwiGLU tile interleave)")
        fs = slice(rank * (ff // tp), (rank + 1) * (ff // tp))
        g1, u1 = ffp[fs], up[fs]
        nb = g1.shape[0] // 128
        w13 = torch.stack([g1.reshape(nb, 128, d), u1.reshape(nb, 128, d)], dim=1).reshape(2 * nb * 128, d)
        out[p + "w13"] = w13
        out[p + "w2"] = g(p + "ff_out.weight")[:, fs]
        out[p + "attn_norm"] = g(p + "attn_norm.weight")
        out[p + "ff_norm"] = g(p + "ff_norm.weight")
    return out


class TensorParallelLLaDA:
    """Same call contract as model.LLaDAForMultiModalGeneration (`forward_rows`, `__call__`), one rank of a TP group."""

    def __init__(self, config, state_dict: Dict[str, torch.Tensor], tp_rank: int, tp_size: int, group=None,
                 max_seq_len: Optional[int] = None, max_batch: int = 1, device: str = "cuda:0", text_vocab_size: int = 126356,
                 codebook_size: int = 8192):
        if not torch.cuda.is_available():
            raise _lib.MmdpError("mmada_parallel_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.config, self.group, self.rank, self.tp = config, group, tp_rank, tp_size
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        g = lambda k, dflt=None: getattr(config, k, dflt)
        self.d_model, self.n_heads, self.n_layers = int(g("d_model")), int(g("n_heads")), int(g("n_layers"))
        self.ff = int(g("mlp_hidden_size") or g("mlp_ratio", 4) * self.d_model)
        self.vocab_rows = int(g("embedding_size") or g("vocab_size"))
        self.rms_eps = float(g("rms_norm_eps", 1e-5))
        self.h_local = self.n_heads // tp_size
        self.d_attn = self.h_local * 128
        if self.d_attn % 256:
            raise ValueError("n_heads / tp must be even (the fused QKV+RoPE epilogue works on 2-head tiles)")
        self.ff_local = self.ff // tp_size
        self.v_local = self.vocab_rows // tp_size
        self.vq_col0, self.vq_cols = text_vocab_size, codebook_size
        self.c_local = codebook_size // tp_size
        self.max_seq_len = int(max_seq_len or g("max_sequence_length", 4096))
        self.max_batch = max_batch
        sh = shard_state_dict(state_dict, self.n_layers, self.n_heads, tp_rank, tp_size, text_vocab_size, codebook_size)
        self.w = {k: v.detach().to(device=self.device, dtype=torch.bfloat16).contiguous() for k, v in sh.items()}
        cos, sin = rope_tables(128, float(g("rope_theta", 10000.0)), self.max_seq_len)
        self.cos, self.sin = cos.to(self.device), sin.to(self.device)
        M, d, bf = self.max_batch * self.max_seq_len, self.d_model, dict(dtype=torch.bfloat16, device=self.device)
        self.x = torch.empty((M, d), **bf)
        self.xn = torch.empty((M, d), **bf)
        self.q = torch.empty((M, self.d_attn), **bf)
        self.k = torch.empty((M, self.d_attn), **bf)
        self.att = torch.empty((M, self.d_attn), **bf)
        self.h = torch.empty((M, self.ff_local), **bf)
        self.part = torch.empty((M, d), dtype=torch.float32, device=self.device)
        self.vt = None
        self._vt_key = None

    def eval(self):
        return self

    def _allreduce(self, t: torch.Tensor):
        if self.tp > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _gather_cols(self, local: torch.Tensor, out: torch.Tensor):
        """local [n, c] on every rank -> out [n, tp*c] (rank r's block at columns r*c)."""
        if self.tp == 1:
            out.copy_(local)
            return
        n, c = local.shape
        buf = torch.empty((self.tp, n, c), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf, local.contiguous(), group=self.group)
        out.view(n, self.tp, c).copy_(buf.permute(1, 0, 2))

    def _hidden(self, ids: torch.Tensor):
        B, L = ids.shape
        M, d, s = B * L, self.d_model, stream_ptr()
        if M > self.x.shape[0]:
            raise _lib.MmdpError("TensorParallelLLaDA: batch x length exceeds the workspace")
        Lpad = (L + 7) // 8 * 8
        if self._vt_key != (B, Lpad):
            self.vt = torch.zeros((B, self.h_local, 128, Lpad), dtype=torch.bfloat16, device=self.device)
            self._vt_key = (B, Lpad)
        x, xn, part, w = self.x[:M], self.xn[:M], self.part[:M], self.w
        check(lib.mmdp_embed(ptr(ids), ptr(w["wte"]), ptr(x), M, d, w["wte"].shape[0], s))
        scale = 1.0 / math.sqrt(128.0)
        for i in range(self.n_layers):
            p = f"blocks.{i}."
            check(lib.mmdp_rmsnorm(ptr(x), d, None, ptr(w[p + "attn_norm"]), ptr(xn), d, M, d, self.rms_eps, s))
            check(lib.mmdp_qkv_rope_tp(ptr(xn), d, ptr(w[p + "wqkv"]), M, d, self.h_local, L, Lpad, ptr(self.cos), ptr(self.sin),
                                       ptr(self.q), ptr(self.k), ptr(self.vt), s))
            check(lib.mmdp_attention(ptr(self.q), ptr(self.k), ptr(self.vt), ptr(self.att), B, self.h_local, L, Lpad, scale, s))
            check(lib.mmdp_gemm_bf16(EPI_F32, ptr(self.att), self.d_attn, ptr(w[p + "wo"]), self.d_attn, M, d, self.d_attn,
                                     ptr(part), d, None, 0, s))
            self._allreduce(part)
            check(lib.mmdp_resid_add_f32(ptr(x), d, ptr(part), d, M, d, s))
            check(lib.mmdp_rmsnorm(ptr(x), d, None, ptr(w[p + "ff_norm"]), ptr(xn), d, M, d, self.rms_eps, s))
            check(lib.mmdp_gemm_bf16(EPI_SWIGLU, ptr(xn), d, ptr(w[p + "w13"]), d, M, 2 * self.ff_local, d, ptr(self.h),
                                     self.ff_local, None, 0, s))
            check(lib.mmdp_gemm_bf16(EPI_F32, ptr(self.h), self.ff_local, ptr(w[p + "w2"]), self.ff_local, M, d, self.ff_local,
                                     ptr(part), d, None, 0, s))
            self._allreduce(part)
            check(lib.mmdp_resid_add_f32(ptr(x), d, ptr(part), d, M, d, s))
        return x

    @torch.no_grad()
    def forward_rows(self, ids: torch.Tensor, rows_a: Optional[torch.Tensor] = None, rows_b: Optional[torch.Tensor] = None,
                     col0_b: int = 0, ncols_b: int = 0, out_a: Optional[torch.Tensor] = None, out_b: Optional[torch.Tensor] = None):
        x = self._hidden(ids.contiguous())
        d, s, w = self.d_model, stream_ptr(), self.w
        ra = rb = None
        if rows_a is not None and rows_a.numel():
            n = rows_a.numel()
            xr = torch.empty((n, d), dtype=torch.bfloat16, device=self.device)
            check(lib.mmdp_rmsnorm(ptr(x), d, ptr(rows_a), ptr(w["ln_f"]), ptr(xr), d, n, d, self.rms_eps, s))
            loc = _lib.gemm_bf16(xr, w["head"], EPI_PLAIN)
            ra = out_a if out_a is not None else torch.empty((n, self.vocab_rows), dtype=torch.bfloat16, device=self.device)
            self._gather_cols(loc, ra)
        if rows_b is not None and rows_b.numel():
            if col0_b != self.vq_col0 or ncols_b != self.vq_cols:
                raise _lib.MmdpError("TensorParallelLLaDA: the column window must be the VQ codebook window given at construction")
            n = rows_b.numel()
            xr = torch.empty((n, d), dtype=torch.bfloat16, device=self.device)
            check(lib.mmdp_rmsnorm(ptr(x), d, ptr(rows_b), ptr(w["ln_f"]), ptr(xr), d, n, d, self.rms_eps, s))
            loc = _lib.gemm_bf16(xr, w["head_vq"], EPI_PLAIN)
            rb = out_b if out_b is not None else torch.empty((n, ncols_b), dtype=torch.bfloat16, device=self.device)
            self._gather_cols(loc, rb)
        return ra, rb

    @torch.no_grad()
    def forward(self, input_ids=None, infer: bool = True, use_cache: bool = False, **_) -> ModelOutput:
        ids = input_ids.to(device=self.device, dtype=torch.int64)
        if ids.dim() == 1:
            ids = ids.unsqueeze(0)
        B, L = ids.shape
        rows = torch.arange(B * L, dtype=torch.int32, device=self.device)
        logits, _ = self.forward_rows(ids.contiguous(), rows_a=rows)
        return ModelOutput(logits=logits.view(B, L, self.vocab_rows))

    __call__ = forward
