"""Sampling-chain kernels (text step, image step A/M, re-mask) against the oracle on identical seeded inputs, through
the C ABI. Integer outputs (token ids, masks, mask_len) must be bit-exact; the fp64 text confidence is compared at
1e-12 relative (summation order). bf16 probabilities may differ from torch-CPU softmax in a ~1e-5 fraction of entries
by one ulp (different expf implementations), so image-step cases are constructed with inputs on which decisions are
checked for equality and any mismatch must be explained by such a one-ulp probability difference."""
import pytest
import torch

from helpers import load_golden
from oracle import sampling as S

pytestmark = pytest.mark.gpu

MASK = 126336


def _lib():
    from mmada_parallel_b200 import _lib
    return _lib


def run_text_step(logits, ids, k, uncond=None, cfg=0.0, unoise=None, temperature=0.0):
    L = _lib()
    dev = "cuda"
    lg = logits.to(dev).contiguous()
    R, V = lg.shape
    un = uncond.to(dev).contiguous() if uncond is not None else None
    nz = unoise.to(dev).contiguous() if unoise is not None else None
    idsd = ids.to(dev).clone()
    x0 = torch.empty(R, dtype=torch.int64, device=dev)
    conf = torch.empty(R, dtype=torch.float64, device=dev)
    L.check(L.lib.mmdp_text_step(L.ptr(lg), L.ptr(un), V, R, V, float(cfg), L.ptr(nz), V, float(temperature), L.ptr(idsd), MASK,
                                 int(k), L.ptr(x0), L.ptr(conf), L.stream_ptr()))
    torch.cuda.synchronize()
    return idsd.cpu(), x0.cpu(), conf.cpu()


@pytest.mark.parametrize("R,V,k,seed", [(16, 4096, 2, 0), (256, 134656, 2, 1), (33, 8192, 40, 2), (8, 1024, 0, 3)])
def test_text_step_greedy(R, V, k, seed):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(R, V, generator=g) * 2).to(torch.bfloat16)
    logits[0, 5] = logits[0, 900] = 30.0  # exact tie: first index wins
    ids = torch.where(torch.rand(R, generator=g) < 0.6, torch.tensor(MASK), torch.randint(0, V, (R,), generator=g))
    new_o, x0_o, conf_o = S.text_step(logits, ids, MASK, k)
    new_g, x0_g, conf_g = run_text_step(logits, ids, k)
    masked = ids == MASK
    assert torch.equal(x0_g[masked], x0_o[masked]) and int(x0_g[0]) == 5
    assert torch.allclose(conf_g, conf_o, rtol=1e-12, atol=0)
    assert torch.equal(new_g, new_o)


def test_text_step_cfg_and_gumbel():
    g = torch.Generator().manual_seed(7)
    R, V = 12, 16384
    cond = (torch.randn(R, V, generator=g) * 2).to(torch.bfloat16)
    unc = (torch.randn(R, V, generator=g) * 2).to(torch.bfloat16)
    ids = torch.full((R,), MASK)
    ids[3] = 11
    new_o, x0_o, conf_o = S.text_step(cond, ids, MASK, 3, uncond_logits=unc, text_cfg=2.5)            # variant M
    new_g, x0_g, conf_g = run_text_step(cond, ids, 3, uncond=unc, cfg=2.5)
    assert torch.equal(new_g, new_o) and torch.allclose(conf_g, conf_o, rtol=1e-12, atol=0)
    u = torch.rand(R, V, generator=g).to(torch.bfloat16)
    new_o, x0_o, conf_o = S.text_step(cond, ids, MASK, 4, temperature=0.7, uniform_noise=u)          # variant A + Gumbel
    new_g, x0_g, conf_g = run_text_step(cond, ids, 4, unoise=u, temperature=0.7)
    # logf differs between libm and CUDA in the last ulp, which bf16 rounding almost always hides; require equality of
    # the decisions and report otherwise
    assert torch.equal(x0_g[ids == MASK], x0_o[ids == MASK])
    assert torch.equal(new_g, new_o) and torch.allclose(conf_g, conf_o, rtol=1e-12, atol=0)


def test_add_gumbel_kat():
    kat = load_golden("sampler_kat.pt")
    for c in kat["a_add_gumbel_noise"]:
        R = c["logits"].shape[0]
        ids = torch.full((R,), MASK)
        _, x0, _ = run_text_step(c["logits"], ids, 0, unoise=c["uniform"], temperature=c["temp"])
        assert torch.equal(x0, c["argmax"])


def run_image_step(variant, cond, ua, ub, s_a, s_b, ids, pos, sched, temp, q, noise, vq_offset, C=8192, want_probs=True):
    L = _lib()
    dev = "cuda"
    N = cond.shape[0]
    d = lambda t: t.to(dev).contiguous() if t is not None else None
    cond, ua, ub, q, noise = d(cond), d(ua), d(ub), d(q), d(noise)
    idsd = ids.to(dev).clone()
    posd = pos.to(dev).to(torch.int32)
    sampled = torch.empty(N, dtype=torch.int32, device=dev)
    selp = torch.empty(N, dtype=torch.float32, device=dev)
    unk = torch.empty(N, dtype=torch.uint8, device=dev)
    probs = torch.empty((N, C), dtype=torch.bfloat16, device=dev)
    mlen = torch.zeros(1, dtype=torch.int32, device=dev)
    masking = torch.empty(N, dtype=torch.uint8, device=dev)
    L.check(L.lib.mmdp_image_step(variant, L.ptr(cond), L.ptr(ua), L.ptr(ub), C, N, C, float(s_a), float(s_b), L.ptr(q),
                                  L.ptr(noise), float(temp), int(sched), L.ptr(idsd), L.ptr(posd), MASK, vq_offset,
                                  L.ptr(sampled), L.ptr(selp), L.ptr(unk), L.ptr(probs), L.ptr(mlen), L.ptr(masking),
                                  L.stream_ptr()))
    torch.cuda.synchronize()
    return dict(ids=idsd.cpu(), sampled=sampled.cpu().long(), selp=selp.cpu(), unknown=unk.cpu().bool(), probs=probs.cpu(),
                mask_len=int(mlen.item()), masking=masking.cpu().bool())


def _image_case(seed, N=1024, frac_known=0.4, C=8192):
    g = torch.Generator().manual_seed(seed)
    cond = (torch.randn(N, C, generator=g) * 2.5).to(torch.bfloat16)
    ua = (torch.randn(N, C, generator=g) * 2.5).to(torch.bfloat16)
    ub = (torch.randn(N, C, generator=g) * 2.5).to(torch.bfloat16)
    known = torch.rand(N, generator=g) < frac_known
    vq = torch.where(known, torch.randint(0, C, (N,), generator=g), torch.tensor(-1))
    q = torch.empty(N, C, dtype=torch.bfloat16).exponential_(1, generator=g)
    rn = torch.randn(N, generator=g).to(torch.bfloat16)
    un = torch.rand(N, generator=g).to(torch.bfloat16)
    return cond, ua, ub, vq, q, rn, un


def _check_against_oracle(out, orc, variant):
    """Decisions must match the oracle; where they do not, the row's probabilities must differ by a bf16 ulp (expf
    implementation), which is reported and bounded."""
    pm = (out["probs"] != orc["probs"])
    assert pm.float().mean() < 1e-3, f"bf16 probs differ in {pm.float().mean():.2e} of entries"
    rows_diff = pm.any(-1)
    same = out["sampled"] == orc["sampled"]
    assert bool(same[~rows_diff].all()), "sampled ids differ on rows whose probabilities are bit-identical"
    assert same.float().mean() > 0.995
    assert out["mask_len"] == orc["mask_len"]
    if bool(same.all()) and torch.equal(out["selp"].to(torch.bfloat16), orc["selected_probs"].to(torch.bfloat16)):
        assert torch.equal(out["masking"], orc["masking"])
    return bool(same.all())


def _assert_writeback_exact(variant, out, ids, pos, sched, temp, noise, OFF, N):
    """UNCONDITIONAL check of the integer stage (confidence -> mask_len -> re-mask -> id write-back): the oracle's
    mask_by_random_topk is evaluated on the kernel's OWN stage-1 results (sampled id, selected bf16 probability, unknown
    flag), so there is no floating-point freedom left - mask_len, the mask and every written id must be equal."""
    unknown = out["unknown"]
    selected = out["selp"].to(torch.bfloat16)
    mask_len = max(1, min(int(unknown.sum()) - 1, sched))
    assert out["mask_len"] == (min(mask_len, N - 1) if variant == "A" else mask_len)
    if variant == "A":
        masking, _ = S.mask_by_random_topk_a(mask_len, selected, temp, noise, stable=True)
    else:
        masking, _ = S.mask_by_random_topk_m(mask_len, selected, temp, noise)
    assert torch.equal(out["masking"], masking), "re-mask differs from the oracle on identical stage-1 results"
    want = ids.clone()
    want[pos] = torch.where(masking, torch.tensor(MASK), out["sampled"] + OFF)
    assert torch.equal(out["ids"], want), "id write-back"


@pytest.mark.parametrize("seed,s_a,s_b,temp,use_q,sched", [(0, 0.0, 4.0, 0.5, True, 600), (1, 1.5, 4.0, 0.0, False, 3),
                                                          (2, 0.0, 0.0, 1.0, True, 2000), (3, 2.0, 0.0, 0.25, True, -1)])
def test_image_step_a(seed, s_a, s_b, temp, use_q, sched):
    cond, ua, ub, vq, q, rn, _ = _image_case(seed)
    N, OFF = cond.shape[0], 126356
    pos = torch.arange(N) + (torch.arange(N) // 32) + 10  # VQ positions with a newline slot every 32 tokens
    ids = torch.zeros(int(pos.max()) + 5, dtype=torch.int64)
    ids[pos] = torch.where(vq == -1, torch.tensor(MASK), vq + OFF)
    orc = S.image_step("A", cond, ua, ub, s_a, s_b, vq, MASK, sched, temp, q if use_q else None, rn, 8192, stable=True)
    out = run_image_step(0, cond, ua if s_a else None, ub if s_b else None, s_a, s_b, ids, pos, sched, temp,
                         q if use_q else None, rn, OFF)
    assert torch.equal(out["unknown"], orc["unknown"])
    _assert_writeback_exact("A", out, ids, pos, sched, temp, rn, OFF, N)
    if _check_against_oracle(out, orc, "A"):
        want_ids = ids.clone()
        want_ids[pos] = torch.where(orc["final"] == -1, torch.tensor(MASK), orc["final"] + OFF)
        if torch.equal(out["masking"], orc["masking"]):
            assert torch.equal(out["ids"], want_ids)
    # tie-independent invariants of the re-mask
    k = out["mask_len"]
    assert int(out["masking"].sum()) == min(k, N - 1)
    assert not bool(out["masking"][~out["unknown"]].any()) or k > int(out["unknown"].sum())


@pytest.mark.parametrize("seed,s,temp,sched", [(10, 4.0, 0.6, 500), (11, 3.5, 0.0, 1), (12, 0.0, 1.0, 900)])
def test_image_step_m(seed, s, temp, sched):
    cond, ua, _, vq, q, _, un = _image_case(seed)
    N, OFF = cond.shape[0], 126349
    vqm = torch.where(vq == -1, torch.tensor(MASK), vq)
    pos = torch.arange(N) + 7
    ids = torch.zeros(N + 20, dtype=torch.int64)
    ids[pos] = torch.where(vq == -1, torch.tensor(MASK), vq + OFF)
    orc = S.image_step("M", cond, ua, None, s, 0.0, vqm, MASK, sched, temp, q, un, 8192)
    out = run_image_step(1, cond, ua, None, s, 1 + s, ids, pos, sched, temp, q, un, OFF)
    assert torch.equal(out["unknown"], orc["unknown"])
    _assert_writeback_exact("M", out, ids, pos, sched, temp, un, OFF, N)
    if _check_against_oracle(out, orc, "M") and torch.equal(out["masking"], orc["masking"]):
        want = ids.clone()
        want[pos] = torch.where(orc["masking"], torch.tensor(MASK), orc["sampled"] + OFF)
        assert torch.equal(out["ids"], want)


def test_remask_known_answers_from_reference():
    """mask_by_random_topk fixtures produced by the REAL reference functions (tests/golden/sampler_kat.pt), through
    mmdp_image_remask. Variant M (cut-off form) must match exactly; variant A must match the reference's mask except
    among tokens whose confidence EQUALS the boundary value (torch.sort's tie order is unspecified), and must match the
    oracle's stable rule exactly."""
    L = _lib()
    kat = load_golden("sampler_kat.pt")
    dev = "cuda"
    for variant, key in ((0, "a_mask_by_random_topk"), (1, "m_mask_by_random_topk")):
        for c in kat[key]:
            N = c["probs"].numel()
            sampled = torch.arange(N, dtype=torch.int32, device=dev)
            selp = c["probs"].float().to(dev)
            unk = torch.ones(N, dtype=torch.uint8, device=dev)
            noise = c["noise"].to(dev).contiguous()
            ids = torch.zeros(N, dtype=torch.int64, device=dev)
            pos = torch.arange(N, dtype=torch.int32, device=dev)
            mlen = torch.zeros(1, dtype=torch.int32, device=dev)
            masking = torch.empty(N, dtype=torch.uint8, device=dev)
            L.check(L.lib.mmdp_image_remask(variant, N, L.ptr(sampled), L.ptr(selp), L.ptr(unk), L.ptr(noise), float(c["temp"]),
                                            int(c["k"]), L.ptr(ids), L.ptr(pos), MASK, 1000, L.ptr(mlen), L.ptr(masking), L.stream_ptr()))
            torch.cuda.synchronize()
            got = masking.cpu().bool()
            ref = c["masking"]
            if variant == 1:
                assert torch.equal(got, ref)
                want_ids = torch.where(ref, torch.tensor(MASK), torch.arange(N) + 1000)
            else:
                stable, conf = S.mask_by_random_topk_a(c["k"], c["probs"], c["temp"], c["noise"], stable=True)
                assert torch.equal(got, stable)
                assert int(got.sum()) == int(ref.sum()) == min(c["k"], N - 1)
                diff = got != ref
                if bool(diff.any()):
                    assert conf[diff].float().unique().numel() == 1, "differs from the reference beyond boundary ties"
                want_ids = torch.where(stable, torch.tensor(MASK), torch.arange(N) + 1000)
            assert torch.equal(ids.cpu(), want_ids)


@pytest.mark.parametrize("variant", ["A", "M"])
def test_image_step_margin_safe_end_to_end(variant):
    """Whole image step against the oracle with NO conditional: inputs are built so that no decision sits within
    floating-point reach of a boundary - every row has one dominant code (logit +12 over unit noise, so its probability is
    ~1 in bf16 on both sides whatever the last bit of expf), greedy sampling, and temperature 0 with distinct dominant
    logits per row is replaced by a confidence noise whose magnitude (temp 4 x randn) separates the positions."""
    g = torch.Generator().manual_seed(123)
    N, C = 256, 8192
    cond = torch.randn(N, C, generator=g).to(torch.bfloat16)
    unc = torch.randn(N, C, generator=g).to(torch.bfloat16)
    win = torch.randint(0, C, (N,), generator=g)
    cond[torch.arange(N), win] = 14.0
    unc[torch.arange(N), win] = 10.0
    vq = torch.where(torch.rand(N, generator=g) < 0.3, torch.randint(0, C, (N,), generator=g), torch.tensor(-1))
    noise = (torch.randn(N, generator=g) if variant == "A" else torch.rand(N, generator=g)).to(torch.bfloat16)
    OFF = 126356 if variant == "A" else 126349
    pos = torch.arange(N) + 5
    ids = torch.zeros(N + 10, dtype=torch.int64)
    ids[pos] = torch.where(vq == -1, torch.tensor(MASK), vq + OFF)
    sched, temp = 77, 4.0
    if variant == "A":
        orc = S.image_step("A", cond, None, unc, 0.0, 2.0, vq, MASK, sched, temp, None, noise, C, stable=True)
        out = run_image_step(0, cond, None, unc, 0.0, 2.0, ids, pos, sched, temp, None, noise, OFF)
        final = orc["final"]
    else:
        vqm = torch.where(vq == -1, torch.tensor(MASK), vq)
        orc = S.image_step("M", cond, unc, None, 2.0, 0.0, vqm, MASK, sched, temp, None, noise, C)
        out = run_image_step(1, cond, unc, None, 2.0, 3.0, ids, pos, sched, temp, None, noise, OFF)
        final = torch.where(orc["masking"], torch.tensor(-1), orc["sampled"])
    # the construction must really be margin-safe: distinct confidences around the cut
    conf = orc["confidence"].float()
    srt = conf.sort().values
    assert float((srt[1:] - srt[:-1])[max(0, orc["mask_len"] - 2): orc["mask_len"] + 1].min()) > 0
    assert torch.equal(out["sampled"], orc["sampled"]) and bool((out["sampled"][orc["unknown"]] == win[orc["unknown"]]).all())
    assert out["mask_len"] == orc["mask_len"] and torch.equal(out["masking"], orc["masking"])
    want = ids.clone()
    want[pos] = torch.where(final == -1, torch.tensor(MASK), final + OFF)
    assert torch.equal(out["ids"], want)


def test_image_and_text_step_large_counts():
    """768x768 images of the reference app are 2304 VQ tokens; text spans above 1024 positions: the single-CTA re-mask and
    commit kernels loop over their elements (limits 4096), results equal the oracle's integer logic."""
    g = torch.Generator().manual_seed(5)
    N, C, OFF = 2304, 8192, 126356
    cond = torch.randn(N, C, generator=g).to(torch.bfloat16)
    win = torch.randint(0, C, (N,), generator=g)
    cond[torch.arange(N), win] = 14.0
    vq = torch.where(torch.rand(N, generator=g) < 0.5, torch.randint(0, C, (N,), generator=g), torch.tensor(-1))
    rn = torch.randn(N, generator=g).to(torch.bfloat16)
    pos = torch.arange(N) + (torch.arange(N) // 48) + 3
    ids = torch.zeros(int(pos.max()) + 4, dtype=torch.int64)
    ids[pos] = torch.where(vq == -1, torch.tensor(MASK), vq + OFF)
    out = run_image_step(0, cond, None, None, 0.0, 0.0, ids, pos, 700, 3.0, None, rn, OFF)
    _assert_writeback_exact("A", out, ids, pos, 700, 3.0, rn, OFF, N)
    R, V = 1500, 4096
    logits = (torch.randn(R, V, generator=g) * 2).to(torch.bfloat16)
    tids = torch.where(torch.rand(R, generator=g) < 0.7, torch.tensor(MASK), torch.randint(0, V, (R,), generator=g))
    new_o, x0_o, conf_o = S.text_step(logits, tids, MASK, 37)
    new_g, x0_g, conf_g = run_text_step(logits, tids, 37)
    assert torch.equal(new_g, new_o) and torch.allclose(conf_g, conf_o, rtol=1e-12, atol=0)
