"""Repeats the full-size one-layer forward and reports bitwise differences between runs (debugging aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmada_parallel_b200.model import LLaDAForMultiModalGeneration
from mmada_parallel_b200 import _lib
from oracle.llada import make_config

cfg = make_config(d_model=4096, n_heads=32, n_layers=1, mlp_hidden_size=12288, vocab_size=134656, max_sequence_length=2432)
m = LLaDAForMultiModalGeneration(cfg, max_seq_len=2432, max_batch=2)
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, std: (torch.randn(*s, device="cuda", generator=g) * std).to(torch.bfloat16)
d, ff, V = 4096, 12288, 134656
sd = {"model.transformer.wte.weight": rnd(V, d, std=0.02), "model.transformer.ff_out.weight": rnd(V, d, std=d ** -0.5),
      "model.transformer.ln_f.weight": torch.ones(d, device="cuda", dtype=torch.bfloat16)}
p = "model.transformer.blocks.0."
for n, shape, std in [("q_proj", (d, d), d ** -0.5), ("k_proj", (d, d), d ** -0.5), ("v_proj", (d, d), d ** -0.5),
                      ("attn_out", (d, d), d ** -0.5), ("ff_proj", (ff, d), d ** -0.5), ("up_proj", (ff, d), d ** -0.5),
                      ("ff_out", (d, ff), ff ** -0.5)]:
    sd[p + n + ".weight"] = rnd(*shape, std=std)
sd[p + "attn_norm.weight"] = torch.ones(d, device="cuda", dtype=torch.bfloat16)
sd[p + "ff_norm.weight"] = torch.ones(d, device="cuda", dtype=torch.bfloat16)
m.load_state_dict(sd)
L = 2414
ids = torch.randint(0, 126000, (2, L), device="cuda", generator=g)
rows = torch.cat([torch.arange(2157, 2413), torch.arange(L + 1100, L + 1100 + 64)]).to(torch.int32).cuda()
for ver in [int(v) for v in os.environ.get('MMDP_DET_VERS', '6,3').split(',')]:
    ref, _ = m.forward_rows(ids, rows_a=rows)
    ref = ref.clone()
    hid = m.hidden_state().clone() if hasattr(m, "hidden_state") else None
    bad = 0
    for it in range(int(os.environ.get('MMDP_DET_ITERS', '12'))):
        out, _ = m.forward_rows(ids, rows_a=rows)
        diff = (out != ref)
        if bool(diff.any()):
            bad += 1
            r = diff.any(dim=1).nonzero().flatten().tolist()
            print(f"attn v{ver} iter {it}: {int(diff.sum())} elements differ in rows {r[:8]} (of {len(r)}), max |d| {float((out.float()-ref.float()).abs().max()):.4f}")
    print(f"attn v{ver}: {bad} runs differ")
