"""Worker for tests/test_gpu_tp.py (launched with torchrun, one process per GPU): tensor-parallel forward vs the
single-GPU forward on the same seeded weights, and rank-consistency of a short TP generation."""
import contextlib
import io
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import llada  # noqa: E402  (tests may use the oracle's seeded weight generator)
from oracle.make_golden import layout_a  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
from mmada_parallel_b200.generators.parallel_generator import generate_ti2ti  # noqa: E402
from mmada_parallel_b200.model import LLaDAForMultiModalGeneration  # noqa: E402
from mmada_parallel_b200.tensor_parallel import TensorParallelLLaDA  # noqa: E402

cfg = llada.make_config(d_model=512, n_heads=4, n_layers=2, mlp_hidden_size=1024, vocab_size=134656,
                        max_sequence_length=512)
sd = llada.make_weights(cfg, seed=77)
tp = TensorParallelLLaDA(cfg, sd, rank, world, max_seq_len=512, device=f"cuda:{rank}")              # NVLink peer-memory collective
tp_nccl = TensorParallelLLaDA(cfg, sd, rank, world, max_seq_len=512, device=f"cuda:{rank}", collective="nccl")
lay = layout_a()
ids = lay["input_ids"].cuda()
lg_tp = tp(ids).logits
lg_nccl = tp_nccl(ids).logits
ok = True
# same rounding points, different fp32 summation order across ranks (fixed rank order vs NCCL's): inside the 4-ulp bound, and
# the peer-memory path is bitwise repeatable
d_modes = (lg_tp.float() - lg_nccl.float()).abs().max().item()
rep = torch.equal(tp(ids).logits, lg_tp)
g = [None] * world
dist.all_gather_object(g, (d_modes, rep, float(lg_tp.float().abs().max())))
if rank == 0:
    print("p2p vs nccl collective: max |dlogit| per rank", [round(x[0], 4) for x in g], "repeatable", [x[1] for x in g])
    ok = ok and all(x[1] for x in g) and all(x[0] <= 4 * x[2] * 2.0 ** -8 for x in g)
# every rank holds identical logits (the activations are broadcast, the head slices all-gathered)
ref = lg_tp.clone()
dist.broadcast(ref, src=0)
same_logits = torch.tensor([1 if torch.equal(ref, lg_tp) else 0], device=f"cuda:{rank}")
dist.all_reduce(same_logits, op=dist.ReduceOp.MIN)
ok = ok and int(same_logits.item()) == 1
if rank == 0:
    single = LLaDAForMultiModalGeneration(cfg, max_seq_len=512, max_batch=1, device="cuda:0")
    single.load_state_dict(sd)
    lg_1 = single(ids, infer=True).logits
    err = (lg_tp.float() - lg_1.float()).abs()
    scale = lg_1.float().abs().max().item()
    tol = 4 * scale * 2.0 ** -8
    agree = (lg_tp.argmax(-1) == lg_1.argmax(-1)).float().mean().item()
    print(f"TP{world} vs single: max err {err.max().item():.4f} mean {err.mean().item():.5f} tol {tol:.4f} argmax agree {agree:.3f}")
    ok = ok and err.max().item() <= tol and agree > 0.9
# two row chunks on two streams (chunk_split: sequences of >= 1024 rows) against the one-chunk schedule: every row's partial sums are
# the same GEMM results reduced in the same rank order, so the logits must be bitwise identical
tp_c2 = TensorParallelLLaDA(cfg, sd, rank, world, max_seq_len=1600, device=f"cuda:{rank}", chunks=2)
tp_c1 = TensorParallelLLaDA(cfg, sd, rank, world, max_seq_len=1600, device=f"cuda:{rank}", chunks=1)
ids_long = torch.randint(0, 126000, (1, 1500), generator=torch.Generator().manual_seed(3)).cuda()
rows = torch.cat([torch.arange(0, 48), torch.arange(740, 800), torch.arange(1450, 1500)]).to(torch.int32).cuda()
la2, _ = tp_c2.forward_rows(ids_long, rows_a=rows)
la1, _ = tp_c1.forward_rows(ids_long, rows_a=rows)
la2b, _ = tp_c2.forward_rows(ids_long, rows_a=rows)
chunk_ok = torch.tensor([1 if (torch.equal(la1, la2) and torch.equal(la2, la2b) and not torch.isnan(la2.float()).any()) else 0], device=f"cuda:{rank}")
dist.all_reduce(chunk_ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("two-chunk pipelined forward bitwise equal to the one-chunk forward, repeatable:", bool(chunk_ok.item()),
          "max |d|", float((la1.float() - la2.float()).abs().max()))
ok = ok and int(chunk_ok.item()) == 1
del tp_c1, tp_c2
args = {k: lay[k] for k in ("text_start", "text_end", "image_start", "seq_len", "newline_every", "uncon_text", "uncon_image")}
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(5)
    img, txt = generate_ti2ti(tp, lay["input_ids"], text_steps=8, timesteps=4, text_gen_length=16, temperature=1.0, text_temperature=0.0,
                              cfg_scale=0.0, cfg_img=4.0, generator=torch.Generator(device=f"cuda:{rank}").manual_seed(42), **args)
t = torch.tensor(img + txt, dtype=torch.int64, device=f"cuda:{rank}")
gathered = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    print("ranks produced identical token sequences:", same)
    print("TP_CHECK_OK" if (ok and same) else "TP_CHECK_FAILED")
dist.destroy_process_group()
