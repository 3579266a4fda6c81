"""GPU parity tests (`-m gpu`): the CUDA path, called through the C-ABI (ctypes), against the CPU oracle on the
same seeded inputs, against the frozen golden vectors, and through size-independent properties at larger sizes.

Tolerances (stated, see DESIGN.md "parity"): integer / index work (bucket search, Philox, compaction) is
bit-exact given identical logits and uniforms; the transformer runs bf16 operands with fp32 accumulation against
the oracle's fp32, so logits are compared with |dlogit| <= LOGIT_ATOL (max) and per-dimension log-prob with
|dlogp| <= LOGP_ATOL.
"""
import math
import os
import pickle

import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_ATOL = 0.20      # max |logit_cuda - logit_oracle|; logits have std ~2.6; measured max 0.09-0.131 on B200 (r2), bar = 1.5 x that;
                       # against the bf16-EMULATING oracle the gap is 0.05 (tests/test_gpu_parity_r2.py::test_bf16_emulated_oracle)
LOGIT_MEAN_ATOL = 0.035  # measured mean 0.018-0.023
LOGP_ATOL = 0.08       # per dimension; measured 0.04
CDF_ATOL = 0.06        # end-to-end shift of the inverse-CDF position in probability mass (buckets / 5000)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.pt")


def _toy(N, dx, dth, seed):
    """theta ~ N(0, I), x = theta W^T + 0.1 eps + 1 (generator of /root/reference/tests/test_npe_pfn.py:47-55)."""
    g = torch.Generator().manual_seed(seed)
    theta = torch.randn(N, dth, generator=g)
    w = torch.randn(dx, dth, generator=g)
    x = theta @ w.T + 0.1 * torch.randn(N, dx, generator=g) + 1.0
    return theta, x, g


def test_library_loaded_and_abi(engine):
    from npe_pfn_b200.engine import ABI_SYMBOLS, load_library
    L = load_library()
    assert L.pfn_abi_version() == 2
    for s in ABI_SYMBOLS:
        assert hasattr(L, s)
    assert engine.launch_count >= 1


@pytest.mark.parametrize("F,N", [(1, 16), (2, 33), (3, 100), (5, 257), (10, 300)])
def test_fit_statistics_and_borders(engine, weights, F, N):
    from oracle import bar_head
    from oracle import tabpfn_oracle as model
    from oracle.estimator import y_standardise
    g = torch.Generator().manual_seed(N)
    X = torch.randn(N, F, generator=g) * 3 + 1
    if F >= 3:
        X[1, 1] = float("nan")
        X[2, 2] = float("inf")
    y = torch.randn(N, generator=g) * 2 - 0.5
    engine.prefill(0, X, y)
    ex = engine.slot_export(0)
    ym, ys, yz = y_standardise(y)
    st = model.EncoderStats(X, yz, (F + 1) // 2)
    assert torch.allclose(ex["mean"].cpu(), st.mean, atol=1e-6, rtol=1e-6)
    assert torch.allclose(ex["std"].cpu(), st.std, atol=1e-6, rtol=1e-6)
    assert torch.equal(ex["scale"].cpu(), st.scale)
    assert abs(float(ex["y_mean"]) - ym) <= 1e-6 * max(1, abs(ym)) and abs(float(ex["y_std"]) - ys) <= 1e-6 * ys
    assert abs(float(ex["y_fill"]) - float(st.y_fill)) < 1e-6
    if float(ex["y_mean"]) == ym and float(ex["y_std"]) == ys:
        assert torch.equal(ex["borders"].cpu(), bar_head.renorm_borders(weights.borders, ym, ys))
    assert ex["N"] == N and ex["F"] == F and ex["T"] == (F + 1) // 2 + 1


@pytest.mark.parametrize("F,N,M", [(2, 1, 3), (3, 2, 1), (1, 16, 8), (2, 64, 50), (3, 100, 33), (5, 200, 130), (10, 300, 64), (19, 150, 40),
                                   (40, 70, 20)])
def test_logits_and_kv_cache_vs_oracle(engine, weights, F, N, M):
    from oracle.estimator import OracleTabPFNRegressor
    g = torch.Generator().manual_seed(F * 1000 + N)
    Xc = torch.randn(N, F, generator=g)
    yc = Xc[:, 0] * 0.8 + 0.2 * torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    if F >= 3:
        Xc[0, 1] = float("nan")
        Xt[0, 2] = float("-inf")
    oracle = OracleTabPFNRegressor(weights=weights).fit(Xc, yc)
    ref = oracle.predict(Xt)["logits"]
    engine.prefill(1, Xc, yc)
    got = engine.forward_logits(1, Xt).cpu()
    d = (got - ref).abs()
    print(f"F={F} N={N}: max|dlogit|={d.max():.4f} mean={d.mean():.5f} ref std={ref.std():.3f}")
    assert d.max() <= LOGIT_ATOL and d.mean() <= LOGIT_MEAN_ATOL
    # K/V cache of the context (bf16) against the oracle's fp32 head-0 K/V
    kv = engine.slot_export(1, want_kv=True)["kv"].float().cpu()  # [L, T, N, 64]
    for l in (0, weights.cfg.nlayers - 1):
        k_ref, v_ref = oracle.cache.k0[l], oracle.cache.v0[l]  # [T, N, 32]
        assert (kv[l, :, :, :32] - k_ref).abs().max() <= 0.08
        assert (kv[l, :, :, 32:] - v_ref).abs().max() <= 0.08


def test_golden_vectors_cuda(engine, weights):
    gold = torch.load(GOLDEN)
    for case in gold["cases"]:
        engine.prefill(2, case["Xc"], case["yc"])
        got = engine.forward_logits(2, case["Xt"]).cpu()
        d = (got[:, case["cols"]] - case["logits_cols"]).abs()
        assert d.max() <= LOGIT_ATOL, (case["F"], d.max())
        assert (torch.logsumexp(got, -1) - case["lse"]).abs().max() <= LOGIT_ATOL
        # head on the GOLDEN logits row: identical bucket and sample, bit for bit
        th, bins, _ = engine.head_sample(2, case["logits_full0"], uniforms=case["u"][:1], return_bins=True)
        assert int(bins[0]) == int(case["idx0"]) and float(th[0]) == float(case["theta0"])


@pytest.mark.parametrize("M,B_scale", [(1, 1.0), (257, 3.0), (4096, 0.3)])
def test_head_bit_exact_vs_oracle(engine, weights, M, B_scale):
    from oracle import bar_head
    g = torch.Generator().manual_seed(M)
    B = weights.cfg.num_buckets
    logits = torch.randn(M, B, generator=g) * B_scale
    if M > 4:
        logits[3, :100] = -float("inf")
        logits[4] = 0.0
    u = torch.rand(M, generator=g)
    yv = torch.randn(200, generator=g)
    engine.prefill(3, torch.randn(200, 2, generator=g), yv * 1.7 + 0.3)
    borders = engine.slot_export(3)["borders"].cpu()
    th, bins, lp = engine.head_sample(3, logits, uniforms=u, return_bins=True, with_log_prob=True)
    th_ref, idx_ref, _ = bar_head.sample(logits, borders, uniforms=u)
    assert torch.equal(bins.cpu(), idx_ref)
    assert torch.equal(th.cpu(), th_ref)
    nll_ref = bar_head.nll(logits, borders, th_ref)
    assert torch.allclose(-lp.cpu(), nll_ref, atol=2e-5, rtol=1e-5)
    # Philox path: same uniforms as the oracle's generator, hence same draws
    th2, bins2, _ = engine.head_sample(3, logits, seed=1234567, row0=10, offset=3, return_bins=True)
    th2_ref, idx2_ref, _ = bar_head.sample(logits, borders, seed=1234567, row0=10, offset=3)
    assert torch.equal(bins2.cpu(), idx2_ref) and torch.equal(th2.cpu(), th2_ref)
    # negative log density on given targets, including both half-normal tails
    y = torch.cat([torch.linspace(borders[0].item() - 2, borders[-1].item() + 2, M - 1), borders[5:6]])[:M] \
        if M > 1 else borders[7:8]
    nll = engine.head_nll(3, logits, y).cpu()
    assert torch.allclose(nll, bar_head.nll(logits, borders, y), atol=2e-5, rtol=1e-5)
    # broadcast of a single logits row to many draws
    th3, bins3, _ = engine.head_sample(3, logits[:1], M=64, uniforms=torch.linspace(0.01, 0.99, 64), return_bins=True)
    th3_ref, idx3_ref, _ = bar_head.sample(logits[:1].expand(64, B).contiguous(), borders,
                                           uniforms=torch.linspace(0.01, 0.99, 64))
    assert torch.equal(bins3.cpu(), idx3_ref) and torch.equal(th3.cpu(), th3_ref)
    assert torch.all(th3[1:] >= th3[:-1])  # inverse CDF is monotone in u


def test_fused_sampling_vs_reference_loop(engine, weights):
    """`_sample` with injected uniforms against the oracle's restatement of npe_pfn.py:111-169."""
    from npe_pfn_b200 import NPE_PFN_Core
    from oracle.estimator import OracleTabPFNRegressor
    from oracle.reference_loop import sample_loop
    theta, x, g = _toy(120, 3, 3, 7)
    xo = x[:1].clone()
    M = 96
    u = torch.rand(M, 3, generator=g)
    post = NPE_PFN_Core(regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    s, lp, bins = post._sample(M, xo, with_log_prob=True, uniforms=u, return_bins=True)
    s_ref, lp_ref, bins_ref = sample_loop(OracleTabPFNRegressor(weights=weights), x, theta, xo, M,
                                          with_log_prob=True, uniforms=u, return_bins=True)
    # Bucket identity holds for identical logits (test_head_bit_exact_vs_oracle).  End to end the bf16 logits move
    # the CDF a little, so the bucket found for the same uniform moves by a few of the 5000 equal-mass buckets:
    # bound that shift in probability mass, and the resulting draw in units of the posterior spread.
    B = weights.cfg.num_buckets
    dbin = (bins.cpu() - bins_ref).abs().float()
    print("|dbucket| median per dim:", dbin.median(0).values.tolist(), "max", dbin.max().item())
    assert dbin[:, 0].max() / B <= CDF_ATOL and dbin[:, 0].median() / B <= CDF_ATOL / 4
    assert dbin.max() / B <= 2 * CDF_ATOL
    spread = s_ref.std(0)
    dth = (s - s_ref).abs() / spread
    print("|dtheta| / posterior std: median", dth.median(0).values.tolist(), "max", dth.max(0).values.tolist())
    assert dth[:, 0].median() <= 0.05 and dth.median() <= 0.1
    assert (s.mean(0) - s_ref.mean(0)).abs().max() <= 0.1 * spread.max()
    # log-density of a draw is only comparable where both paths landed in the same buckets (the random-init
    # density differs by e^2.6 between neighbouring buckets)
    same = (bins.cpu() == bins_ref).all(1)
    assert same.any()
    assert (lp - lp_ref).abs()[same].max() <= 3 * LOGP_ATOL


def test_log_prob_vs_reference_loop(engine, weights):
    from npe_pfn_b200 import NPE_PFN_Core
    from oracle.estimator import OracleTabPFNRegressor
    from oracle.reference_loop import logprob_loop
    theta, x, g = _toy(150, 4, 3, 9)
    xo = x[:1].clone()
    th = torch.randn(80, 3, generator=g)
    th[0] = 50.0  # far outside: half-normal tail / clamp path
    post = NPE_PFN_Core(regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    lp = post.log_prob(th, xo)
    lp_ref = logprob_loop(OracleTabPFNRegressor(weights=weights), x, theta, xo, th)
    d = (lp - lp_ref).abs()
    print(f"max |dlogp| = {d.max():.4f} (3 dims)")
    assert lp.shape == (80,) and torch.isfinite(lp).all()
    assert d.max() <= 3 * LOGP_ATOL
    # chunked evaluation (max_sampling_batch_size) gives the same numbers
    lp_chunked = post.log_prob(th, xo, max_sampling_batch_size=17)
    assert torch.allclose(lp, lp_chunked, atol=1e-5)


def test_slot_cache_not_shared_between_posteriors(engine):
    """Two posterior objects on one engine with different simulations must never reuse each other's K/V caches
    (regression: slot tags once used id(self), which Python reuses after garbage collection)."""
    from npe_pfn_b200 import NPE_PFN_Core
    theta_a, x_a, g = _toy(60, 2, 2, 100)
    theta_b, x_b, _ = _toy(60, 2, 2, 200)
    th = torch.randn(16, 2, generator=g)
    ref = {}
    for name, (t, x) in {"a": (theta_a, x_a), "b": (theta_b, x_b)}.items():
        ref[name] = NPE_PFN_Core(regressor_init_kwargs={"engine": engine}).append_simulations(t, x).log_prob(th, x[:1])
    assert (ref["a"] - ref["b"]).abs().max() > 1e-3
    for _ in range(3):  # interleave short-lived objects: same context version, possibly recycled ids
        for name, (t, x) in {"a": (theta_a, x_a), "b": (theta_b, x_b)}.items():
            lp = NPE_PFN_Core(regressor_init_kwargs={"engine": engine}).append_simulations(t, x).log_prob(th, x[:1])
            assert torch.equal(lp, ref[name])


@pytest.mark.parametrize("Ntot,dx,k", [(50, 2, 50), (1000, 3, 100), (100_000, 10, 10_000), (1_000_000, 5, 10_000)])
def test_context_filter_on_device(engine, Ntot, dx, k):
    """pfn_filter_context against the reference's torch formula (support_posterior.py:357-369)."""
    from npe_pfn_b200.support_posterior import standardized_euclidean_filtering
    g = torch.Generator().manual_seed(Ntot)
    x = torch.randn(Ntot, dx, generator=g) * torch.arange(1, dx + 1) + 0.5
    theta = torch.arange(Ntot, dtype=torch.float32)[:, None]
    obs = x[Ntot // 3:Ntot // 3 + 1] + 0.01
    idx, dist = engine.filter_context(x, obs[0], k, want_dist=True)
    idx, dist = idx.cpu(), dist.cpu()
    th_ref, x_ref = standardized_euclidean_filtering(obs, theta, x, k)
    idx_ref = th_ref[:, 0].long()
    assert idx.shape == (k,) and len(set(idx.tolist())) == k
    assert bool((dist[1:] >= dist[:-1]).all())  # ascending distance
    mu, sd = x.mean(0), x.std(0)
    d_all = torch.norm((x - mu) / sd - (obs - mu) / sd, dim=1)
    assert torch.allclose(dist, d_all[idx], rtol=1e-4, atol=1e-5)
    # same set up to near-ties at the cut-off, same order up to fp32 rounding of the distances
    common = len(set(idx.tolist()) & set(idx_ref.tolist()))
    assert common >= k - max(2, k // 1000)
    assert float(d_all[idx].max()) <= float(d_all[idx_ref].max()) * (1 + 1e-4) + 1e-6
    agree = (idx == idx_ref).float().mean().item()
    assert agree >= 0.99 or k == Ntot


def test_get_context_uses_device_filter(engine):
    from npe_pfn_b200 import TabPFN_Based_NPE_PFN
    theta, x, g = _toy(300, 3, 2, 12)
    post = TabPFN_Based_NPE_PFN(filter_context_size=40, regressor_init_kwargs={"engine": engine})
    post.append_simulations(theta, x)
    th_d, x_d = post.get_context(x[5])
    post.device_filter = False
    th_h, x_h = post.get_context(x[5])
    assert th_d.shape == (40, 2) and torch.equal(x_d[0], x[5])
    assert torch.equal(th_d, th_h) and torch.equal(x_d, x_h)


def test_empty_inputs(engine):
    """Zero test rows / zero candidates are valid calls through the C-ABI and return empty results."""
    g = torch.Generator().manual_seed(4)
    engine.prefill(9, torch.randn(20, 3, generator=g), torch.randn(20, generator=g))
    assert engine.forward_logits(9, torch.empty(0, 3)).shape[0] == 0
    joint = torch.empty(0, 5, device="cuda")
    engine.sample_step(9, joint, 3, 3, seed=1)
    lp = torch.empty(0, device="cuda")
    engine.logprob_step(9, joint, 3, 3, lp)
    th, bins, _ = engine.head_sample(9, torch.empty(0, 5000, device="cuda"), return_bins=True)
    assert th.shape == (0,) and bins.shape == (0,)
    with pytest.raises(RuntimeError):
        engine.forward_logits(10, torch.randn(2, 3))  # slot never prefilled
    with pytest.raises(RuntimeError):
        engine.prefill(9, torch.randn(4, 200), torch.randn(4))  # more features than the model supports


def test_slot_pack_unpack_roundtrip(engine):
    """A slot exported with slot_pack and installed into another slot with slot_unpack (what ranks exchange in the
    sharded prefill) answers exactly like the original."""
    g = torch.Generator().manual_seed(77)
    Xc, yc, Xt = torch.randn(90, 5, generator=g), torch.randn(90, generator=g), torch.randn(40, 5, generator=g)
    engine.prefill(7, Xc, yc)
    a = engine.forward_logits(7, Xt)
    enc, borders, kv = engine.slot_pack(7)
    engine.slot_unpack(8, 90, 5, enc, borders, kv)
    b = engine.forward_logits(8, Xt)
    assert torch.equal(a, b)
    u = torch.rand(40, generator=g)
    ta, _, _ = engine.head_sample(7, a, uniforms=u)
    tb, _, _ = engine.head_sample(8, b, uniforms=u)
    assert torch.equal(ta, tb) and engine.slot_info(8) == engine.slot_info(7)


def test_accept_compact_matches_torch(engine):
    g = torch.Generator().manual_seed(3)
    for M, dim in [(1, 2), (255, 3), (256, 1), (100_003, 5)]:
        th = (torch.rand(M, dim, generator=g) * 4 - 2).cuda()
        if M > 10:
            th[7, 0] = float("nan")
        lo = torch.full((dim,), -1.0)
        hi = torch.full((dim,), 1.5)
        idx, rows, count = engine.accept_compact(th, lo=lo, hi=hi)
        ok = ((th >= lo.cuda()) & (th <= hi.cuda())).all(1) & torch.isfinite(th).all(1)
        k = int(count)
        assert k == int(ok.sum())
        assert torch.equal(idx[:k], torch.nonzero(ok).squeeze(1))
        assert torch.equal(rows[:k], th[ok])
    # mask-only form, empty input and all-rejected
    th = torch.randn(1000, 2, generator=g).cuda()
    mask = (torch.arange(1000) % 3 == 0)
    idx, rows, count = engine.accept_compact(th, mask=mask)
    assert int(count) == int(mask.sum()) and torch.equal(rows[:int(count)], th[mask.cuda()])
    idx, rows, count = engine.accept_compact(th, lo=torch.full((2,), 100.0), hi=torch.full((2,), 101.0))
    assert int(count) == 0
    idx, rows, count = engine.accept_compact(torch.empty(0, 2, device="cuda"))
    assert int(count) == 0


def test_rows_independent_and_chunk_invariant(engine):
    """Size-independent properties at a larger context: permuting / re-chunking test rows does not change a row."""
    g = torch.Generator().manual_seed(21)
    N, F, M = 2000, 7, 3000
    Xc = torch.randn(N, F, generator=g)
    yc = Xc.sum(1) + 0.1 * torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    engine.prefill(4, Xc, yc)
    a = engine.forward_logits(4, Xt)
    perm = torch.randperm(M, generator=g)
    b = engine.forward_logits(4, Xt[perm])
    assert torch.equal(a[perm.cuda()], b)
    engine.set_option("chunk_rows", 1000)
    try:
        c = engine.forward_logits(4, Xt)
    finally:
        engine.set_option("chunk_rows", 0)
    assert torch.equal(a, c)
    assert torch.isfinite(a).all()


def test_full_size_properties(engine):
    """BASELINE config 2 sizes (10-D theta / 10-D x, 10 000 simulations, 20 000 draws): properties that do not need
    the (slow) oracle - determinism under a seed, invariance to the chunking of test rows, and the round trip
    sample -> log_prob: the log density accumulated while sampling equals log_prob of the same draws."""
    from npe_pfn_b200 import NPE_PFN_Core
    g = torch.Generator().manual_seed(2024)
    N, d, S = 10_000, 10, 20_000
    theta = math.sqrt(0.1) * torch.randn(N, d, generator=g)
    x = theta + math.sqrt(0.1) * torch.randn(N, d, generator=g)
    xo = x[:1].clone()
    prior = torch.distributions.MultivariateNormal(torch.zeros(d), 0.1 * torch.eye(d))
    post = NPE_PFN_Core(prior=prior, regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    s1, lp1 = post._sample(S, xo, with_log_prob=True, seed=123)
    s2, lp2 = post._sample(S, xo, with_log_prob=True, seed=123)
    assert torch.equal(s1, s2) and torch.equal(lp1, lp2)
    assert s1.shape == (S, d) and torch.isfinite(s1).all() and torch.isfinite(lp1).all()
    engine.set_option("chunk_rows", 6000)
    try:
        s3, lp3 = post._sample(S, xo, with_log_prob=True, seed=123)
    finally:
        engine.set_option("chunk_rows", 0)
    assert torch.equal(s1, s3) and torch.equal(lp1, lp3)
    s4, _ = post._sample(S, xo, seed=124)
    assert not torch.equal(s1, s4)
    lp = post.log_prob(s1, xo, max_sampling_batch_size=S)
    assert (lp - lp1).abs().max() <= 1e-3
    # draws live inside the bucket range spanned by the simulations (borders are renormalised per dimension)
    assert float(s1.abs().max()) < 10.0
    assert post.engine.slot_info(9)["N"] == N and post.engine.slot_info(9)["T"] == 11


def test_item_attention_v6_agrees(engine, weights):
    """attn_tc6 (row sum on the tensor core, two query tiles per CTA; opt-in `attn_impl = 2`, slower than v5 and therefore
    not the default - profiles/r2_attn_tc6_experiments.txt) against v5 and against the fp32 oracle, ragged sizes included."""
    from oracle.estimator import OracleTabPFNRegressor
    g = torch.Generator().manual_seed(66)
    for (N, F, M) in ((1500, 5, 700), (70, 3, 129), (64, 2, 1)):
        Xc = torch.randn(N, F, generator=g)
        yc = Xc[:, 1] + 0.1 * torch.randn(N, generator=g)
        Xt = torch.randn(M, F, generator=g)
        outs = []
        for impl in (1, 2):
            engine.set_option("attn_impl", impl)
            engine.prefill(5, Xc, yc)
            outs.append(engine.forward_logits(5, Xt).cpu())
        engine.set_option("attn_impl", 1)
        ref = OracleTabPFNRegressor(weights=weights).fit(Xc, yc).predict(Xt)["logits"]
        d56, d6 = (outs[0] - outs[1]).abs(), (outs[1] - ref).abs()
        print(f"N={N}: v5 vs v6 max {d56.max():.4f} mean {d56.mean():.5f} | v6 vs fp32 oracle max {d6.max():.4f} mean {d6.mean():.5f}")
        assert torch.isfinite(outs[1]).all()
        assert d6.max() <= LOGIT_ATOL and d6.mean() <= LOGIT_MEAN_ATOL and d56.mean() <= LOGIT_MEAN_ATOL


def test_item_attention_impls_agree(engine):
    """tcgen05 item attention against the warp-level mma.sync implementation (both bf16 in, fp32 accumulate)."""
    g = torch.Generator().manual_seed(33)
    N, F, M = 1500, 5, 700
    Xc = torch.randn(N, F, generator=g)
    yc = Xc[:, 1] + 0.1 * torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    outs = []
    for impl in (0, 1):
        engine.set_option("attn_impl", impl)
        engine.prefill(5, Xc, yc)
        outs.append(engine.forward_logits(5, Xt))
    engine.set_option("attn_impl", 1)
    d = (outs[0] - outs[1]).abs()
    print(f"mma vs tcgen05 item attention: max|dlogit|={d.max():.4f}")
    assert d.max() <= LOGIT_ATOL


def test_gemm_impls_agree(engine):
    """tcgen05 projections (fused LayerNorm / GELU / bias epilogues) against the warp-level mma.sync GEMMs."""
    g = torch.Generator().manual_seed(44)
    N, F, M = 700, 6, 900
    Xc = torch.randn(N, F, generator=g)
    yc = Xc[:, 2] - Xc[:, 0] + 0.1 * torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    outs = []
    for impl in (0, 1):
        engine.set_option("gemm_impl", impl)
        engine.prefill(5, Xc, yc)
        outs.append(engine.forward_logits(5, Xt))
    engine.set_option("gemm_impl", 1)
    d = (outs[0] - outs[1]).abs()
    print(f"mma vs tcgen05 GEMMs: max|dlogit|={d.max():.4f} mean={d.mean():.5f}")
    assert torch.isfinite(outs[1]).all()
    assert d.max() <= LOGIT_ATOL and d.mean() <= LOGIT_MEAN_ATOL


# ---- reference-facing API (shapes / errors as in /root/reference/tests/test_npe_pfn.py) -------------------
def test_api_sample_logprob_shapes_and_errors(engine):
    from npe_pfn_b200 import BoxUniform, TabPFN_Based_NPE_PFN
    theta, x, g = _toy(50, 2, 2, 1)
    prior = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    for filt in ("standardized_euclidean_filtering", "latest_filtering", "random_filtering", "no_filtering"):
        post = TabPFN_Based_NPE_PFN(prior=prior, filter_type=filt, filter_context_size=30,
                                    regressor_init_kwargs={"engine": engine})
        post.append_simulations(theta, x)
        s = post.sample((30,), x[0])
        assert s.shape == (30, 2) and s.device.type == "cpu" and torch.isfinite(s).all()
        s2, lp2 = post.sample((10,), x[:1], with_log_prob=True)
        assert s2.shape == (10, 2) and lp2.shape == (10,) and torch.isfinite(lp2).all()
        lp = post.log_prob(s, x[0])
        assert lp.shape == (30,) and torch.isfinite(lp).all()
    with pytest.raises(ValueError):
        post.sample((5,), x[:2])
    with pytest.raises(ValueError):
        post.log_prob(s, x[0], mode="nope")
    with pytest.raises(ValueError):
        TabPFN_Based_NPE_PFN(filter_type="unknown")
    # box prior: real rejection, all draws inside the support, batch-size adaptation exercised
    box = BoxUniform(-0.3 * torch.ones(2), 0.8 * torch.ones(2))
    post = TabPFN_Based_NPE_PFN(prior=box, regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    s = post.sample((500,), x[0], max_sampling_batch_size=200)
    assert s.shape == (500, 2) and bool(box.support.check(s).all())
    assert 0 < post.last_acceptance_rate <= 1
    # elementwise-support prior (plain Uniform, as in the reference demo) takes the torch.all(dim=-1) branch
    uni = torch.distributions.Uniform(-torch.ones(2), torch.ones(2))
    post = TabPFN_Based_NPE_PFN(prior=uni, regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    s = post.sample((64,), x[0])
    assert s.shape == (64, 2) and bool(((s >= -1) & (s <= 1)).all())
    # pickling drops the engine-backed model and rebuilds it
    blob = pickle.dumps(NPE_PFN_CoreFactory(engine, theta, x))
    post2 = pickle.loads(blob)
    assert post2.sample((3,), x[0]).shape == (3, 2)


def NPE_PFN_CoreFactory(engine, theta, x):
    from npe_pfn_b200 import NPE_PFN_Core
    prior = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    return NPE_PFN_Core(prior=prior).append_simulations(theta, x)


def test_api_sample_batched(engine):
    from npe_pfn_b200 import BoxUniform, NPE_PFN_Core
    theta, x, g = _toy(60, 3, 2, 2)
    post = NPE_PFN_Core(prior=None, regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    s = post.sample_batched(x[:5], (20,))
    assert s.shape == (5, 20, 2) and torch.isfinite(s).all()
    s, lp = post.sample_batched(x[:3], (8,), with_log_prob=True)
    assert s.shape == (3, 8, 2) and lp.shape == (3, 8)
    box = BoxUniform(-2 * torch.ones(2), 2 * torch.ones(2))
    post = NPE_PFN_Core(prior=box, regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    s, lp = post.sample_batched(x[:4], (25,), with_log_prob=True)
    assert s.shape == (4, 25, 2) and lp.shape == (4, 25) and bool(box.support.check(s.reshape(-1, 2)).all())
    # a single observation through sample_batched has the shape of sample (test_npe_pfn.py:361-382)
    assert post.sample_batched(x[:1], (10,)).shape == (1, 10, 2)
    # the same observation repeated gives per-observation draws from the same distribution
    s = post.sample_batched(x[:1].repeat(2, 1), (400,))
    assert (s[0].mean(0) - s[1].mean(0)).abs().max() < 0.5


def test_five_call_protocol(engine, weights):
    """fit / predict / criterion.sample / criterion(logits, y), as npe_pfn.py:140-151 calls them."""
    from npe_pfn_b200.estimator import B200TabPFNRegressor
    theta, x, g = _toy(40, 2, 1, 4)
    m = B200TabPFNRegressor(engine=engine, slot=6)
    m.fit(x, theta[:, 0])
    pd = m.predict(x[:9], output_type="full", quantiles=[])
    assert set(pd) >= {"criterion", "logits"} and pd["logits"].shape == (9, weights.cfg.num_buckets)
    torch.manual_seed(0)
    a = pd["criterion"].sample(pd["logits"])
    torch.manual_seed(0)
    b = pd["criterion"].sample(pd["logits"])
    assert a.shape == (9,) and a.device.type == "cpu" and torch.equal(a, b)
    nll = pd["criterion"](pd["logits"], a)
    assert nll.shape == (9,) and torch.isfinite(nll).all()


def test_classifier_and_ratio_log_prob(engine, weights):
    """TabPFN classifier head (npe_pfn.py:610, 661, 697) and the ratio-based log-prob built on it (:526-570)."""
    from npe_pfn_b200 import BoxUniform, TabPFN_Based_NPE_PFN
    from npe_pfn_b200.estimator import B200TabPFNClassifier, default_classifier_weights
    from oracle.classifier import OracleTabPFNClassifier
    g = torch.Generator().manual_seed(8)
    n, d = 150, 3
    X = torch.cat([torch.rand(n, d, generator=g) * 4 - 2, torch.randn(n, d, generator=g) * 0.5], 0)
    y = torch.cat([torch.zeros(n), torch.ones(n)])
    Xt = torch.randn(64, d, generator=g)
    got = B200TabPFNClassifier().fit(X, y).predict_proba(Xt)
    ref = OracleTabPFNClassifier(weights=default_classifier_weights()).fit(X, y).predict_proba(Xt)
    assert got.shape == (64, 2) and abs(got.sum(1) - 1).max() < 1e-5
    print("classifier max |dp| =", abs(got - ref).max())
    assert abs(got - ref).max() <= 0.02
    # ratio-based log-prob through the posterior object: finite, cached classifier is reused, bounds exposed
    theta, x, _ = _toy(80, 2, 2, 3)
    prior = BoxUniform(-4 * torch.ones(2), 4 * torch.ones(2))
    post = TabPFN_Based_NPE_PFN(prior=prior, regressor_init_kwargs={"engine": engine}).append_simulations(theta, x)
    th = torch.randn(40, 2, generator=g)
    th[0] = 100.0  # outside the classifier's box -> floor value
    lp1 = post.log_prob(th, x[:1], mode="ratio_based", num_posterior_samples=200)
    assert lp1.shape == (40,) and torch.isfinite(lp1).all()
    lo, hi = post._get_classifier_bounds()
    assert lo.shape == (2,) and bool((hi > lo).all())
    wrapper = post._model_classifier
    assert not wrapper.refit_necessary(x[:1], *reversed(post.get_context(x[:1])), 200, 0.1)
    lp2 = post.log_prob(th, x[:1], mode="ratio_based", num_posterior_samples=200)
    assert torch.equal(lp1, lp2)  # same fitted classifier, deterministic prediction
    floor = float(wrapper._uniform_log_prob + math.log(1e-15) - math.log(1 + 1e-15))
    assert abs(float(lp1[0]) - floor) < 1e-3 and float(lp1.min()) >= floor - 1e-3  # outside the box -> floor value
    assert wrapper.refit_necessary(x[1:2], *reversed(post.get_context(x[1:2])), 200, 0.1)


@pytest.mark.parametrize("num_clusters", [1, 3])
def test_unconditional_estimator(engine, num_clusters):
    """TabPFN_Based_Uncond_Estimator (npe_pfn.py:747-900; shapes / finiteness as tests/test_npe_pfn.py:292-317)."""
    from npe_pfn_b200 import TabPFN_Based_Uncond_Estimator
    torch.manual_seed(0)
    theta = torch.cat([torch.randn(60, 2) * 0.3 + c for c in (-2.0, 0.0, 2.0)])
    est = TabPFN_Based_Uncond_Estimator(num_clusters=num_clusters, regressor_init_kwargs={"engine": engine})
    est.append_simulations(theta)
    s = est.sample((50,))
    assert s.shape == (50, 2) and torch.isfinite(s).all()
    s2, lp2 = est.sample((20,), with_log_prob=True)
    assert s2.shape == (20, 2) and lp2.shape == (20,) and torch.isfinite(lp2).all()
    lp = est.log_prob(s)
    assert lp.shape == (50,) and torch.isfinite(lp).all()
    assert est.cluster_state == 0
    with pytest.raises(ValueError):
        est.log_prob(s, mode="bogus")


def test_tsnpe_rounds_ratio_based(engine):
    """run_tsnpe_pfn with its default log_prob_mode ("ratio_based", tsnpe_pfn.py:25): classifier bounds drive
    prereject_with_bounds in the support proposal."""
    from npe_pfn_b200 import BoxUniform, run_tsnpe_pfn
    torch.manual_seed(1)
    prior = BoxUniform(-2 * torch.ones(2), 2 * torch.ones(2))
    post = run_tsnpe_pfn(lambda t: t + 0.1 * torch.randn_like(t), prior, torch.zeros(1, 2), num_simulations=90,
                         num_rounds=2, proposal_batch_size=100, simulation_batch_size=45,
                         num_samples_to_estimate_support=100, allowed_false_negatives=0.05,
                         regressor_init_kwargs={"engine": engine})
    assert post._theta_train.shape == (90, 2) and bool(prior.support.check(post._theta_train).all())


def test_tsnpe_rounds_autoregressive(engine):
    from npe_pfn_b200 import BoxUniform, run_tsnpe_pfn
    torch.manual_seed(0)
    prior = BoxUniform(-2 * torch.ones(2), 2 * torch.ones(2))

    def simulator(theta):
        return theta + 0.1 * torch.randn_like(theta)

    post = run_tsnpe_pfn(simulator, prior, torch.zeros(1, 2), num_simulations=120, num_rounds=2,
                         proposal_batch_size=200, simulation_batch_size=60, num_samples_to_estimate_support=200,
                         allowed_false_negatives=0.01, log_prob_mode="autoregressive",
                         regressor_init_kwargs={"engine": engine})
    assert post._theta_train.shape == (120, 2)
    s = post.sample((50,), torch.zeros(1, 2))
    assert s.shape == (50, 2) and bool(prior.support.check(s).all())


@pytest.mark.parametrize("gain,nlayers", [(2.5, 12), (6.0, 2), (25.0, 2)])
def test_item_attention_sharp_scores(weights, gain, nlayers):
    """Item attention with LARGE, sharply peaked scores (Q/K projections scaled up): the running maximum of a row keeps
    growing by more than 2^8 along the keys, so the tcgen05 kernel's reference-change machinery runs for real -- v5's
    overflow check + redo / re-reference, the stale tiles after a change, v4's lazy rescaling -- and must agree with the
    warp-level mma.sync kernel (exact running maximum per tile) and with the fp32 oracle.  The two larger gains use a
    2-layer model: with 12 layers of near-argmax attention the network is chaotic (any rounding difference grows to
    ~0.4 in the logits, for every implementation), which would hide what this test is after."""
    import dataclasses
    from npe_pfn_b200.engine import Engine
    from npe_pfn_b200.weights import PFNWeights
    from oracle.estimator import OracleTabPFNRegressor
    base_w = weights if nlayers == weights.cfg.nlayers else PFNWeights.random_init(dataclasses.replace(weights.cfg, nlayers=nlayers))
    t = {k: v.clone() for k, v in base_w.t.items()}
    t["item_wqkv"][:, :2 * base_w.cfg.emsize] *= gain  # Q and K rows: scores grow by gain^2
    w = PFNWeights(base_w.cfg, t)
    eng = Engine(weights=w, max_slots=2)
    g = torch.Generator().manual_seed(int(gain))
    N, F, M = 1300, 3, 300
    Xc = torch.randn(N, F, generator=g)
    Xc = Xc[torch.argsort(Xc[:, 0])]  # ordered context: scores trend along the key index
    yc = Xc[:, 0] + 0.1 * torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    outs = {}
    for name, opts in {"mma": {"attn_impl": 0}, "v4": {"attn_impl": 1, "attn_lean": 0}, "v5": {"attn_impl": 1, "attn_lean": 1},
                       "v5_mufu": {"attn_impl": 1, "attn_lean": 1, "attn_poly": 0},
                       "v5b": {"attn_impl": 1, "attn_lean": 2}}.items():
        eng.set_option("attn_poly", 5)
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.set_option("attn_debug", 1)  # zeroes the event counters
        eng.prefill(0, Xc, yc)
        outs[name] = eng.forward_logits(0, Xt).cpu()
        assert torch.isfinite(outs[name]).all(), name
        if name.startswith("v"):
            redo, changes, general = eng.attn_debug_counts()
            print(f"gain {gain} {name}: redone tiles {redo}, reference changes {changes}, general-path tiles {general}")
            assert changes > 0, "the test must exercise reference-maximum changes"
            if name in ("v5", "v5_mufu"):
                assert redo > 0, "the test must exercise the overflow check + redo path"
    ref = OracleTabPFNRegressor(weights=w).fit(Xc, yc).predict(Xt)["logits"]
    err = {}
    for name, o in outs.items():
        d = (o - ref).abs()
        err[name] = (float(d.max()), float(d.mean()))
        print(f"gain {gain} L={nlayers} {name}: max|dlogit|={d.max():.4f} mean={d.mean():.5f}")
    # every tensor-core variant is as close to the fp32 oracle as the mma.sync kernel is (same bf16 operands)
    for name in ("v4", "v5", "v5_mufu", "v5b"):
        assert err[name][0] <= max(2.0 * err["mma"][0], LOGIT_ATOL), (name, err)
        assert err[name][1] <= max(1.5 * err["mma"][1], LOGIT_MEAN_ATOL), (name, err)
    # and to the mma.sync kernel itself
    d4 = (outs["v4"] - outs["mma"]).abs().mean()
    for name in ("v5", "v5b"):
        dn = (outs[name] - outs["mma"]).abs().mean()
        print(f"gain {gain}: v4 vs mma {d4:.5f} | {name} vs mma {dn:.5f}")
        assert dn <= 1.5 * d4 + 0.01, name
    eng.close()


@pytest.mark.parametrize("F,N,M", [(2, 70, 129), (5, 300, 500), (10, 130, 257), (19, 90, 1000), (29, 64, 77), (30, 40, 1)])
def test_fused_qkv_feature_attention_bit_equal(engine, F, N, M):
    """gemm_tc EPI_FEATURE_ATTN (QKV projection + attention between the tokens of a row in one kernel, qkv never written
    to HBM) against the two-kernel path: the same bf16 q / k / v, the same mma.sync arithmetic -> the SAME bits, for the
    context pass (K/V cache) and for the test rows (logits), T = 2 .. 16 tokens per row incl. rows that straddle nothing
    (an M tile holds floor(128 / T) whole rows) and ragged last tiles."""
    g = torch.Generator().manual_seed(F * 31 + N)
    Xc = torch.randn(N, F, generator=g)
    yc = Xc[:, 0] + 0.1 * torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    outs, kvs = [], []
    for fused in (0, 1):
        engine.set_option("feat_fused", fused)
        engine.prefill(6, Xc, yc)
        outs.append(engine.forward_logits(6, Xt))
        kvs.append(engine.slot_export(6, want_kv=True)["kv"])
    engine.set_option("feat_fused", 1)
    assert torch.isfinite(outs[1]).all()
    assert torch.equal(kvs[0], kvs[1])
    assert torch.equal(outs[0], outs[1])


def test_fused_mlp_kernel_agrees(engine):
    """The fused MLP kernel (mlp_tc.cuh: up-projection, GELU, down-projection, residual, LayerNorm in one launch) against
    the two-kernel tcgen05 path: same bf16 operands and fp32 accumulation, the hidden activation is rounded to bf16 in
    both, so logits agree to rounding noise."""
    g = torch.Generator().manual_seed(55)
    outs = {}
    for (N, F, M) in ((700, 6, 900), (130, 3, 257)):
        Xc = torch.randn(N, F, generator=g)
        yc = Xc[:, 2] - Xc[:, 0] + 0.1 * torch.randn(N, generator=g)
        Xt = torch.randn(M, F, generator=g)
        for fused in (0, 1):
            engine.set_option("mlp_fused", fused)
            engine.prefill(6, Xc, yc)
            outs[fused] = engine.forward_logits(6, Xt)
        engine.set_option("mlp_fused", 1)
        d = (outs[0] - outs[1]).abs()
        print(f"fused vs two-kernel MLP (N={N}): max|dlogit|={d.max():.4f} mean={d.mean():.5f}")
        assert torch.isfinite(outs[1]).all()
        assert d.max() <= 0.1 and d.mean() <= 0.01
