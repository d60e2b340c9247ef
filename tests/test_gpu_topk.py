"""K4 parity: fused similarity + top-k + threshold + intervals vs the numpy oracle (oracle/phase1_ref.py):
bit-exact indices (including tie order), counts and float64 intervals; scores within fp32 dot-product rounding."""
import numpy as np
import pytest
import torch

from oracle import phase1_ref as R

pytestmark = pytest.mark.gpu


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _check(model, img, txt, k, thr, ts=None, dur=30.0, vdur=0.0, dtype=torch.float32, base=0):
    img_t = torch.from_numpy(img).cuda().to(dtype)
    s, i, iv, c = model.sim_topk(img_t, torch.from_numpy(txt).cuda(), k, thr,
                                 None if ts is None else torch.from_numpy(ts), base, dur, vdur)
    s, i, iv, c = s.cpu().numpy(), i.cpu().numpy(), iv.cpu().numpy(), c.cpu().numpy()
    ref_scores = img_t.float().cpu().numpy().astype(np.float32) @ txt.T.astype(np.float32)
    for q in range(txt.shape[0]):
        col = ref_scores[:, q]
        # the kernel's own fp32 scores decide the order; compare against the oracle applied to those scores
        dense = model.similarity(img_t, torch.from_numpy(txt).cuda())[:, q].cpu().numpy()
        assert np.abs(dense - col).max() < 2e-5
        want = R.topk_indices(dense, k) if len(dense) else np.array([], int)
        # np.argsort is not a stable descending sort for ties: define ties -> higher index first explicitly
        want = np.lexsort((np.arange(len(dense)), dense))[::-1][:k]
        n_valid = min(k, len(dense))
        assert np.array_equal(i[q, :n_valid] - base, want), (i[q], want)
        assert np.all(i[q, n_valid:] == -1)
        assert np.array_equal(s[q, :n_valid], dense[want])
        assert c[q] == int((dense[want] >= thr).sum())
        for r in range(n_valid):
            t = float(ts[i[q, r]]) if ts is not None else float(i[q, r])
            assert tuple(iv[q, r]) == R.clip_interval(t, dur, vdur if vdur > 0 else None)
    return s, i, iv, c


def test_single_query_planted_winners(model_b32):
    rng = np.random.default_rng(0)
    img = _unit(rng.standard_normal((5000, 512)).astype(np.float32))
    txt = _unit((img[[17, 4242, 999]] * np.array([[3.0], [2.0], [1.5]])).sum(0, keepdims=True)
                + 0.05 * rng.standard_normal((1, 512)).astype(np.float32)).astype(np.float32)
    s, i, _, _ = _check(model_b32, img, txt, 5, 0.25, ts=np.arange(5000) / 25.0)
    assert list(i[0, :3]) == [17, 4242, 999]


@pytest.mark.parametrize("n,q,k", [(63, 1, 5), (63, 1, 15), (3, 1, 5), (1, 1, 1), (1000, 3, 10), (4097, 16, 32),
                                   (2500, 20, 7), (100000, 2, 10)])
def test_shapes_thresholds_and_tails(model_b32, n, q, k):
    rng = np.random.default_rng(n + q)
    img = _unit(rng.standard_normal((n, 512)).astype(np.float32))
    txt = _unit(rng.standard_normal((q, 512)).astype(np.float32))
    _check(model_b32, img, txt, k, 0.05, ts=np.arange(n) * 0.04, vdur=n * 0.04)
    _check(model_b32, img, txt, k, -1.0)


def test_ties_prefer_higher_index(model_b32):
    rng = np.random.default_rng(5)
    base = _unit(rng.standard_normal((40, 512)).astype(np.float32))
    img = np.concatenate([base, base, base])          # every score appears three times
    txt = _unit(rng.standard_normal((2, 512)).astype(np.float32))
    s, i, _, _ = _check(model_b32, img, txt, 9, -1.0)
    assert i[0, 0] > i[0, 1] > i[0, 2] and s[0, 0] == s[0, 1] == s[0, 2]


def test_bf16_cache_and_index_base_and_768(model_b32):
    rng = np.random.default_rng(9)
    img = _unit(rng.standard_normal((3000, 768)).astype(np.float32))
    txt = _unit(rng.standard_normal((4, 768)).astype(np.float32))
    _check(model_b32, img, txt, 10, 0.0, dtype=torch.bfloat16)
    _check(model_b32, img, txt, 10, 0.0, ts=np.arange(10_000) / 30.0, base=7000)


def test_empty_shard(model_b32):
    txt = _unit(np.random.default_rng(1).standard_normal((2, 512)).astype(np.float32))
    s, i, iv, c = model_b32.sim_topk(torch.empty(0, 512, device="cuda"), torch.from_numpy(txt).cuda(), 5, 0.0)
    assert (i.cpu().numpy() == -1).all() and (c.cpu().numpy() == 0).all()


def test_interval_clamps(model_b32):
    """clip_extractor.py:94-111 edge cases: start clamp at 0, end clamp at duration, start beyond duration."""
    img = np.eye(8, 512, dtype=np.float32)
    txt = np.eye(1, 512, dtype=np.float32) * 1.0
    txt[0, :8] = np.linspace(1.0, 0.3, 8)
    txt = _unit(txt)
    ts = np.array([2.0, 14.9, 15.0, 50.0, 95.0, 99.0, 120.0, 300.0])
    _check(model_b32, img, txt, 8, -1.0, ts=ts, dur=30.0, vdur=100.0)
    _check(model_b32, img, txt, 8, -1.0, ts=ts, dur=0.0, vdur=100.0)      # end <= start -> start + 5
    _check(model_b32, img, txt, 8, -1.0, ts=ts, dur=30.0, vdur=0.0)


def test_multi_shard_merge_equals_global(model_b32):
    """Per-shard top-k with global indices -> b200clip_topk_merge == top-k over the whole matrix (the multi-GPU
    path emulated on one GPU: shards are processed one after the other)."""
    rng = np.random.default_rng(3)
    n, q, k, g = 10007, 5, 10, 8
    img = _unit(rng.standard_normal((n, 512)).astype(np.float32))
    img[5000:5004] = img[100]                          # cross-shard ties
    txt = _unit(rng.standard_normal((q, 512)).astype(np.float32))
    from b200clip.distributed import shard_range

    img_t, txt_t = torch.from_numpy(img).cuda(), torch.from_numpy(txt).cuda()
    ts = torch.arange(n, dtype=torch.float64) / 24.0
    cs, ci = [], []
    for r in range(g):
        lo, hi = shard_range(n, r, g)
        s, i, _, _ = model_b32.sim_topk(img_t[lo:hi], txt_t, k, -1.0, ts, index_base=lo)
        cs.append(s)
        ci.append(i)
    ms, mi, miv, mc = model_b32.topk_merge(torch.stack(cs), torch.stack(ci), 0.1, ts, 30.0, n / 24.0)
    ws, wi, wiv, wc = model_b32.sim_topk(img_t, txt_t, k, 0.1, ts, 0, 30.0, n / 24.0)
    assert torch.equal(mi, wi) and torch.equal(ms, ws) and torch.equal(miv, wiv) and torch.equal(mc, wc)
    for qq in range(q):
        o_s, o_i = R.merge_topk_lists(torch.stack(cs)[:, qq].cpu().numpy(), torch.stack(ci)[:, qq].cpu().numpy(), k)
        assert np.array_equal(o_i, mi[qq].cpu().numpy())


def test_tensor_core_path_fp32_cache_sliced(model_b32):
    """An fp32 cache with >= 16 queries is cast to bf16 slice by slice (2^18 rows) and scored on the tensor cores; the
    lists of all slices meet in one merge.  Planted winners (gap >> bf16 rounding) must come back exactly, in order,
    with global indices, from every slice."""
    n, q, k, e = (1 << 18) * 2 + 5000, 24, 5, 512
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randn(n, e, device="cuda", generator=g)
    img = img / img.norm(dim=-1, keepdim=True)
    txt = torch.randn(q, e, device="cuda", generator=g)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    rows = torch.tensor([7, (1 << 18) - 100, (1 << 18) + 100, (1 << 18) + 12345, (1 << 19) + 4900], device="cuda")
    for qq in range(q):
        for r, row in enumerate(rows):          # winner r of query qq: cosine 0.9 - 0.05 r
            c = 0.9 - 0.05 * r
            noise = img[row] - (img[row] @ txt[qq]) * txt[qq]
            img[(row + qq) % n] = c * txt[qq] + (1 - c * c) ** 0.5 * noise / noise.norm()
    s, i, iv, cnt = model_b32.sim_topk(img, txt, k, 0.5, torch.arange(n, dtype=torch.float64), 0, 30.0, float(n))
    want = torch.stack([(rows + qq) % n for qq in range(q)])
    assert torch.equal(i, want)
    assert float((s - torch.tensor([0.9 - 0.05 * r for r in range(5)], device="cuda")).abs().max()) < 1e-2
    assert torch.equal(cnt, torch.full((q,), 5, dtype=torch.int32, device="cuda"))


@pytest.mark.parametrize("n,q,k,e", [(20000, 256, 5, 512), (9000, 40, 8, 768), (4096, 8, 1, 512), (70001, 300, 5, 512),
                                     (150001, 130, 3, 512), (4200, 9, 8, 64)])
def test_tensor_core_path_order_is_exact(model_b32, n, q, k, e):
    """bf16 cache + many queries -> sim_topk_tc_kernel (tcgen05 similarity, top-k fused in the epilogue).  The order
    must equal an argsort (ties -> higher index) of the scores that very kernel computed, and those scores must be
    within 1e-2 of the fp32 product."""
    from b200clip import capi

    rng = np.random.default_rng(n + q)
    img = _unit(rng.standard_normal((n, e)).astype(np.float32))
    img[n // 2: n // 2 + 3] = img[11]                       # exact ties across tiles
    txt = _unit(rng.standard_normal((q, e)).astype(np.float32))
    img_t = torch.from_numpy(img).cuda().bfloat16()
    txt_t = torch.from_numpy(txt).cuda()
    s = torch.empty(q, k, device="cuda")
    i = torch.empty(q, k, device="cuda", dtype=torch.int64)
    c = torch.empty(q, device="cuda", dtype=torch.int32)
    dense = torch.full((n, q), float("nan"), device="cuda")
    model_b32.handle.call("b200clip_sim_topk_dense", capi._p(img_t), capi.BF16, n, e, capi._p(txt_t), q, k, 0.05,
                          capi._p(s), capi._p(i), capi._p(c), capi._p(dense), model_b32._stream())
    d = dense.cpu().numpy()
    assert not np.isnan(d).any()
    ref = img_t.float().cpu().numpy() @ txt.T
    assert np.abs(d - ref).max() < 1e-2
    want = np.stack([np.lexsort((np.arange(n), d[:, j]))[::-1][:k] for j in range(q)])
    assert np.array_equal(i.cpu().numpy(), want)
    assert np.array_equal(s.cpu().numpy(), np.take_along_axis(d.T, want, 1))
    assert np.array_equal(c.cpu().numpy(), (np.take_along_axis(d.T, want, 1) >= 0.05).sum(1))
    # the public entry selects the same k rows, then RE-SCORES them with the fp32 query in the streaming kernel's exact
    # arithmetic and re-sorts: confidences (and threshold decisions) must not depend on how many queries share a batch
    s2, i2, iv2, c2 = model_b32.sim_topk(img_t, txt_t, k, 0.05, torch.arange(n, dtype=torch.float64), 0, 30.0, float(n))
    sd = model_b32.similarity(img_t, txt_t).cpu().numpy()            # streaming kernel, fp32 query
    i_np, i2_np, s2_np = i.cpu().numpy(), i2.cpu().numpy(), s2.cpu().numpy()
    for j in range(q):
        sel = i_np[j]
        order = np.lexsort((sel, sd[sel, j]))[::-1]
        assert np.array_equal(i2_np[j], sel[order]), (j, i2_np[j], sel[order])
        assert np.array_equal(s2_np[j], sd[sel[order], j])
        for r in range(k):
            assert tuple(iv2[j, r].cpu().numpy()) == R.clip_interval(float(i2_np[j, r]), 30.0, float(n))
    assert np.array_equal(c2.cpu().numpy(), (s2_np >= 0.05).sum(1))
    # ... and equal what a single-query call (streaming path) reports for the same rows
    s1, i1, _, _ = model_b32.sim_topk(img_t, txt_t[:1], k, 0.05, torch.arange(n, dtype=torch.float64), 0, 30.0, float(n))
    common = np.intersect1d(i1[0].cpu().numpy(), i2_np[0])
    assert len(common) >= k - 1                                          # bf16 pre-selection may swap the k-th on a near tie
    for row in common:
        assert s1[0][(i1[0] == int(row))].item() == s2[0][(i2[0] == int(row))].item()


@pytest.mark.parametrize("n,q,k", [(5000, 1, 50), (5000, 3, 100), (40, 1, 64), (100000, 2, 33), (1000, 1, 1000),
                                   (3600, 1, 200), (70, 17, 65)])
def test_top_k_beyond_32(model_b32, n, q, k):
    """phase1_mvp.py:145 accepts any top_k: k > 32 is served by ceil(k / 32) passes, each admitting only rows that come
    strictly after the last slot of the previous pass -- the concatenation must be the plain descending argsort,
    including tie order, the -1 padding when k > n, counts and intervals."""
    rng = np.random.default_rng(n * 7 + k)
    img = _unit(rng.standard_normal((n, 512)).astype(np.float32))
    img[n // 3: n // 3 + 5] = img[1]                          # ties that straddle a pass boundary somewhere
    txt = _unit(rng.standard_normal((q, 512)).astype(np.float32))
    _check(model_b32, img, txt, k, 0.02, ts=np.arange(n) * 0.5, vdur=n * 0.5)
    _check(model_b32, img, txt, k, -1.0, dtype=torch.bfloat16, base=11)


def test_all_equal_scores_across_passes(model_b32):
    img = np.tile(_unit(np.ones((1, 512), np.float32)), (100, 1))
    txt = _unit(np.ones((1, 512), np.float32))
    s, i, _, c = _check(model_b32, img, txt, 70, 0.5)
    assert list(i[0]) == list(range(99, 29, -1)) and c[0] == 70


def test_multi_shard_merge_beyond_32(model_b32):
    rng = np.random.default_rng(12)
    n, q, k, g = 6001, 3, 80, 4
    img = _unit(rng.standard_normal((n, 512)).astype(np.float32))
    img[3000:3003] = img[10]
    txt = _unit(rng.standard_normal((q, 512)).astype(np.float32))
    from b200clip.distributed import shard_range

    img_t, txt_t = torch.from_numpy(img).cuda(), torch.from_numpy(txt).cuda()
    ts = torch.arange(n, dtype=torch.float64) / 24.0
    cs, ci = [], []
    for r in range(g):
        lo, hi = shard_range(n, r, g)
        s, i, _, _ = model_b32.sim_topk(img_t[lo:hi], txt_t, k, -1.0, ts, index_base=lo)
        cs.append(s)
        ci.append(i)
    ms, mi, miv, mc = model_b32.topk_merge(torch.stack(cs), torch.stack(ci), 0.05, ts, 30.0, n / 24.0)
    ws, wi, wiv, wc = model_b32.sim_topk(img_t, txt_t, k, 0.05, ts, 0, 30.0, n / 24.0)
    assert torch.equal(mi, wi) and torch.equal(ms, ws) and torch.equal(miv, wiv) and torch.equal(mc, wc)
