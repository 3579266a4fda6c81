"""The UNMODIFIED reference package (`/root/reference/npe_pfn`) driven over the CPU oracle estimator through the
`oracle/shims` stand-ins for `tabpfn` / `sbi`, compared with the restated loops of `oracle/reference_loop.py`.

This pins the loop restatement (feature slicing, target column, autoregressive append, -inf clamp, summed log-prob)
and the rejection / batching semantics against the reference's own code.  It can only run where `/root/reference`
exists (this container); on the GPU box it is skipped.  CPU only."""
import importlib
import os
import sys

import pytest
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "npe_pfn")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_pkg():
    shims = os.path.join(ROOT, "oracle", "shims")
    saved = list(sys.path)
    sys.path.insert(0, shims)
    sys.path.insert(0, REF)
    for m in [k for k in sys.modules if k == "npe_pfn" or k.startswith("npe_pfn.") or k in ("tabpfn", "sbi")]:
        del sys.modules[m]
    try:
        pkg = importlib.import_module("npe_pfn")
        core = importlib.import_module("npe_pfn.npe_pfn")
        yield pkg, core
    finally:
        sys.path[:] = saved
        for m in [k for k in sys.modules if k == "npe_pfn" or k.startswith("npe_pfn.") or k in ("tabpfn", "sbi")
                  or k.startswith("sbi.")]:
            del sys.modules[m]


def _toy(N, dx, dth, seed):
    g = torch.Generator().manual_seed(seed)
    theta = torch.randn(N, dth, generator=g)
    x = theta @ torch.randn(dth, dx, generator=g) + 0.1 * torch.randn(N, dx, generator=g) + 1.0
    return theta, x, g


def test_reference_sample_and_log_prob_equal_restated_loops(ref_pkg, weights):
    pkg, core = ref_pkg
    from oracle.estimator import OracleTabPFNRegressor
    from oracle.reference_loop import logprob_loop, sample_loop
    theta, x, g = _toy(24, 2, 2, 3)
    xo = x[:1].clone()
    post = core.NPE_PFN_Core(prior=torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2)))
    assert type(post._model).__name__ == "OracleTabPFNRegressor"  # the shim is what the reference constructed
    post.append_simulations(theta, x)
    torch.manual_seed(11)
    s_ref, lp_ref = post._sample(9, xo, with_log_prob=True)          # /root/reference/npe_pfn/npe_pfn.py:111-169
    torch.manual_seed(11)
    s_me, lp_me = sample_loop(OracleTabPFNRegressor(weights=weights), x, theta, xo, 9, with_log_prob=True)
    assert torch.equal(s_ref, s_me) and torch.allclose(lp_ref, lp_me, atol=1e-6)
    th = torch.randn(7, 2, generator=g)
    lp_ref = post.log_prob(th, xo)                                   # :412-455 -> :462-524
    lp_me = logprob_loop(OracleTabPFNRegressor(weights=weights), x, theta, xo, th)
    assert torch.allclose(lp_ref, lp_me, atol=1e-6)


def test_reference_public_api_shapes_over_oracle(ref_pkg):
    """The reference's own smoke assertions (tests/test_npe_pfn.py:69-71, 347-358) hold over the oracle."""
    pkg, core = ref_pkg
    theta, x, g = _toy(20, 3, 2, 5)
    prior = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    post = pkg.TabPFN_Based_NPE_PFN(prior=prior, filter_context_size=15)
    post.append_simulations(theta, x)
    s = post.sample((12,), x[0])
    assert s.shape == (12, 2) and torch.isfinite(s).all()
    lp = post.log_prob(s, x[0])
    assert lp.shape == (12,) and torch.isfinite(lp).all()
    sb = core.NPE_PFN_Core(prior=None).append_simulations(theta, x).sample_batched(x[:3], (4,))
    assert sb.shape == (3, 4, 2)
    with pytest.raises(ValueError):
        post.sample((2,), x[:2])


def test_reference_density_ratio_wrapper_equals_mirror(ref_pkg):
    """`DensityRatioWrapper` of the reference (npe_pfn.py:603-704) against the mirror in npe_pfn_b200, both over the
    oracle classifier and the same posterior draws / torch seed: identical box, cache decisions and log-probs."""
    pkg, core = ref_pkg
    from npe_pfn_b200.npe_pfn import DensityRatioWrapper as Mirror
    from oracle.classifier import OracleTabPFNClassifier

    class MirrorOverOracle(Mirror):
        classifier_cls = OracleTabPFNClassifier

    g = torch.Generator().manual_seed(21)
    draws = torch.randn(60, 2, generator=g) * torch.tensor([0.5, 2.0]) + 1.0
    x, xc, tc = torch.randn(1, 3, generator=g), torch.randn(30, 3, generator=g), torch.randn(30, 2, generator=g)
    th = torch.randn(25, 2, generator=g) * 2 + 1.0
    th[0] = 50.0  # outside the padded box
    ref, mine = core.DensityRatioWrapper(), MirrorOverOracle()
    assert ref.refit_necessary(x, xc, tc, 60, 0.1) and mine.refit_necessary(x, xc, tc, 60, 0.1)
    torch.manual_seed(5)
    ref.fit(x, draws, 0.1, xc, tc)
    torch.manual_seed(5)
    mine.fit(x, draws, 0.1, xc, tc)
    assert torch.equal(ref._padded_dim_min, mine._padded_dim_min) and torch.equal(ref._padded_dim_max, mine._padded_dim_max)
    assert torch.equal(ref._uniform_log_prob, mine._uniform_log_prob)
    for args in [(x, xc, tc, 60, 0.1), (x + 1, xc, tc, 60, 0.1), (x, xc[:-1], tc[:-1], 60, 0.1), (x, xc, tc, 61, 0.1),
                 (x, xc, tc, 60, 0.2)]:
        assert ref.refit_necessary(*args) == mine.refit_necessary(*args)
    a, b = ref.ratio_log_probs(th, 1e-15), mine.ratio_log_probs(th, 1e-15)
    assert torch.allclose(a, b, atol=1e-6) and torch.isfinite(a).all()


def test_reference_loops_over_the_ensemble_oracle(ref_pkg, weights):
    """The unmodified reference with `regressor_init_kwargs={"n_estimators": 2}` (upstream's ensemble switch) over the
    ensemble oracle equals the restated loops over the same oracle: the five-call protocol carries the ensemble
    unchanged (logits = log of the member-averaged probabilities, criterion on the common borders)."""
    pkg, core = ref_pkg
    from oracle.ensemble import OracleEnsembleRegressor
    from oracle.reference_loop import logprob_loop, sample_loop
    theta, x, g = _toy(40, 2, 2, 5)
    xo = x[:1].clone()
    kw = {"n_estimators": 2, "weights": weights, "random_state": 4}
    post = core.NPE_PFN_Core(prior=torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2)),
                             regressor_init_kwargs=kw)
    assert type(post._model).__name__ == "OracleEnsembleRegressor"
    post.append_simulations(theta, x)
    torch.manual_seed(3)
    s_ref, lp_ref = post._sample(6, xo, with_log_prob=True)
    torch.manual_seed(3)
    s_me, lp_me = sample_loop(OracleEnsembleRegressor(**kw), x, theta, xo, 6, with_log_prob=True)
    assert torch.equal(s_ref, s_me) and torch.allclose(lp_ref, lp_me, atol=1e-6)
    th = torch.randn(5, 2, generator=g)
    assert torch.allclose(post.log_prob(th, xo), logprob_loop(OracleEnsembleRegressor(**kw), x, theta, xo, th), atol=1e-6)


# ---- round 2: pins of the support loop and of sample_batched against the reference's own classes -----------------------
def _box(lo, hi, d=2):
    from sbi.utils import BoxUniform  # the shim (oracle/shims/sbi), on sys.path through the ref_pkg fixture
    return BoxUniform(lo * torch.ones(d), hi * torch.ones(d))


def test_reference_sample_batched_equals_mirror(ref_pkg, weights):
    """`sample_batched` with a prior (npe_pfn.py:362-410: oversample 1.5x, up to 10 rounds, per observation the first
    in-support draws) of the UNMODIFIED reference over the oracle, against the mirror's cumulative-sum scatter
    (`npe_pfn_b200.npe_pfn.take_first_n`) fed by the oracle's restatement of `_sample_batched` (:171-251): identical
    draws and log-probs, observation by observation, over several rejection rounds."""
    pkg, core = ref_pkg
    from npe_pfn_b200 import NPE_PFN_Core
    from oracle.estimator import OracleTabPFNRegressor
    from oracle.reference_loop import sample_batched_loop
    theta, x, g = _toy(24, 2, 2, 13)
    prior = _box(-0.4, 1.1)  # tight box around part of the posterior mass: most observations need several rounds
    calls = []

    class MirrorOverOracle(NPE_PFN_Core):
        def _sample_batched(self, xs, n, with_log_prob=False, eps=1e-15, return_device=False, **kw):
            calls.append((xs.shape[0], n))
            return sample_batched_loop(OracleTabPFNRegressor(weights=weights), self._x_train, self._theta_train, xs, n,
                                       with_log_prob=with_log_prob, eps=eps)

    ref = core.NPE_PFN_Core(prior=prior).append_simulations(theta, x)
    mir = MirrorOverOracle(prior=prior, regressor_init_kwargs={"n_estimators": 1}).append_simulations(theta, x)
    mir.redraw_all_observations = True  # the reference's literal schedule, so the torch RNG is consumed identically
    xs = x[:4]
    torch.manual_seed(7)
    a, alp = ref.sample_batched(xs, (9,), with_log_prob=True)
    torch.manual_seed(7)
    b, blp = mir.sample_batched(xs, (9,), with_log_prob=True)
    assert len(calls) >= 2, "the case must need more than one rejection round"
    assert a.shape == (4, 9, 2) and torch.equal(a, b) and torch.allclose(alp, blp, atol=1e-6)
    assert bool(prior.support.check(b.reshape(-1, 2)).all())
    torch.manual_seed(8)
    a = ref.sample_batched(xs, (5,))
    torch.manual_seed(8)
    b = mir.sample_batched(xs, (5,))
    assert torch.equal(a, b)
    # default schedule (later rounds draw only for the observations that are still short): same first-round content,
    # every draw inside the support, fewer proposals
    calls.clear()
    mir.redraw_all_observations = False
    torch.manual_seed(7)
    c = mir.sample_batched(xs, (9,))
    assert c.shape == (4, 9, 2) and bool(prior.support.check(c.reshape(-1, 2)).all())
    assert calls[0] == (4, 13) and all(k <= 4 for k, _ in calls[1:])
    # without a prior: one unfiltered pass (npe_pfn.py:351-358)
    torch.manual_seed(3)
    a = core.NPE_PFN_Core(prior=None).append_simulations(theta, x).sample_batched(xs, (3,))
    torch.manual_seed(3)
    b = MirrorOverOracle(prior=None, regressor_init_kwargs={"n_estimators": 1}).append_simulations(theta, x).sample_batched(xs, (3,))
    assert torch.equal(a, b)


@pytest.mark.parametrize("mode_kwargs", [{"mode": "autoregressive"}, {"mode": "ratio_based", "num_posterior_samples": 30}])
def test_reference_posterior_support_rejection_equals_mirror(ref_pkg, mode_kwargs):
    """`PosteriorSupport` in rejection mode (support_posterior.py:13-69, 97-182; with the ratio-based log-prob also
    `prereject_with_bounds`, :264-309): the reference's class and the mirror, both wrapped around the SAME reference
    posterior object over the oracle and started from the same torch seed, give the same threshold, the same accepted
    proposals in the same order, the same acceptance rate, and the same prior top-up when `max_iter` runs out."""
    pkg, core = ref_pkg
    import npe_pfn.support_posterior as ref_sp
    from npe_pfn_b200.support_posterior import PosteriorSupport as Mirror
    theta, x, g = _toy(24, 2, 2, 17)
    prior = _box(-2.5, 2.5)
    obs = x[:1].clone()
    results = {}
    for name, cls in (("ref", ref_sp.PosteriorSupport), ("mirror", Mirror)):
        posterior = pkg.TabPFN_Based_NPE_PFN(prior=prior, filter_type="no_filtering").append_simulations(theta, x)
        torch.manual_seed(1)
        ps = cls(prior, posterior, obs, num_samples_to_estimate_support=30, batch_size_for_estimate_support=30,
                 allowed_false_negatives=0.2, max_iter_rejection=6, log_prob_kwargs=mode_kwargs)
        torch.manual_seed(2)
        s, rate = ps.sample((20,), show_progress_bars=False, sampling_batch_size=16, return_acceptance_rate=True)
        # an (almost) unreachable threshold: the iteration budget runs out and raw prior draws fill the remainder
        ps.thr = ps.thr + 50.0
        ps.max_iter = 2
        torch.manual_seed(3)
        s2 = ps.sample((6,), show_progress_bars=False, sampling_batch_size=8)
        results[name] = (ps.thr, s, rate, s2)
    (thr_a, s_a, r_a, s2_a), (thr_b, s_b, r_b, s2_b) = results["ref"], results["mirror"]
    assert torch.equal(thr_a, thr_b)
    assert s_a.shape == (20, 2) and torch.equal(s_a, s_b) and r_a == pytest.approx(r_b)
    assert s2_a.shape == (6, 2) and torch.equal(s2_a, s2_b)


def test_reference_posterior_support_sir_equals_mirror(ref_pkg):
    """Sampling-importance-resampling mode (support_posterior.py:184-258): groups of `oversample_sir` posterior draws,
    adaptive log-prob quantile, one categorical pick per group, effective sample sizes."""
    pkg, core = ref_pkg
    import npe_pfn.support_posterior as ref_sp
    from npe_pfn_b200.support_posterior import PosteriorSupport as Mirror
    theta, x, g = _toy(24, 2, 2, 19)
    prior = _box(-3.0, 3.0)
    obs = x[:1].clone()
    out = {}
    for name, cls in (("ref", ref_sp.PosteriorSupport), ("mirror", Mirror)):
        posterior = pkg.TabPFN_Based_NPE_PFN(prior=prior, filter_type="no_filtering").append_simulations(theta, x)
        ps = cls(prior, posterior, obs, sampling_method="sir", oversample_sir=4, allowed_false_negatives=0.1)
        torch.manual_seed(5)
        out[name] = ps.sample((7,), show_progress_bars=False, sampling_batch_size=12, return_ess=True)
    (s_a, ess_a), (s_b, ess_b) = out["ref"], out["mirror"]
    assert s_a.shape == (7, 2) and torch.equal(s_a, s_b)
    assert torch.allclose(ess_a, ess_b, atol=1e-6) and bool((ess_b >= 1 - 1e-5).all()) and bool((ess_b <= 4 + 1e-5).all())
    assert bool(prior.support.check(s_b).all())


def test_reference_tsnpe_rounds_equal_mirror(ref_pkg):
    """`run_tsnpe_pfn` (tsnpe_pfn.py:14-119) of the reference against the mirror's round driver, both building the
    REFERENCE's posterior class over the oracle: same simulations round by round (the proposals, the support thresholds
    and the simulator noise all come from the same torch RNG stream)."""
    pkg, core = ref_pkg
    import npe_pfn.tsnpe_pfn as ref_ts
    import npe_pfn_b200.tsnpe_pfn as mir_ts
    prior = _box(-2.0, 2.0)

    def simulator(t):
        return t + 0.1 * torch.randn_like(t)

    kw = dict(num_simulations=36, num_rounds=3, proposal_batch_size=24, simulation_batch_size=12,
              num_samples_to_estimate_support=24, allowed_false_negatives=0.1, log_prob_mode="autoregressive",
              max_iter_rejection=5)
    torch.manual_seed(4)
    a = ref_ts.run_tsnpe_pfn(simulator, prior, torch.zeros(1, 2), **kw)
    saved = mir_ts.TabPFN_Based_NPE_PFN
    mir_ts.TabPFN_Based_NPE_PFN = pkg.TabPFN_Based_NPE_PFN  # the mirror's driver around the reference's posterior class
    try:
        torch.manual_seed(4)
        b = mir_ts.run_tsnpe_pfn(simulator, prior, torch.zeros(1, 2), **kw)
    finally:
        mir_ts.TabPFN_Based_NPE_PFN = saved
    assert a._theta_train.shape == (36, 2)
    assert torch.equal(a._theta_train, b._theta_train) and torch.equal(a._x_train, b._x_train)


def test_reference_uncond_estimator_equals_mirror(ref_pkg, weights):
    """`TabPFN_Based_Uncond_Estimator` (npe_pfn.py:747-900, experimental in the reference): the mirror's cluster logic
    (shuffle, dummy x, k-means, multinomial split, per-cluster context, final permutation, mixture log-prob) against the
    reference's class, both running the REFERENCE's own hot loops over the oracle estimator, same numpy / torch seeds."""
    import numpy as np
    pkg, core = ref_pkg
    from npe_pfn_b200.uncond import TabPFN_Based_Uncond_Estimator as Mirror
    from oracle.estimator import OracleTabPFNRegressor

    class MirrorOverOracle(Mirror):  # the mirror's own methods around the reference's loops and the oracle model
        _sample = core.NPE_PFN_Core._sample
        _autoregressive_log_prob = core.NPE_PFN_Core._autoregressive_log_prob

    g = torch.Generator().manual_seed(3)
    theta = torch.cat([torch.randn(20, 2, generator=g) * 0.3 + c for c in (-2.0, 2.0)])
    out = {}
    for name, cls in (("ref", core.TabPFN_Based_Uncond_Estimator), ("mirror", MirrorOverOracle)):
        est = cls(num_clusters=2, regressor_init_kwargs={"n_estimators": 1})
        est._model = OracleTabPFNRegressor(weights=weights)
        torch.manual_seed(11)
        np.random.seed(11)
        est.append_simulations(theta)  # KMeans(random_state=None) draws from numpy's global generator, seeded above
        s, lp = est.sample((9,), with_log_prob=True)
        lp2 = est.log_prob(s)
        out[name] = (est._theta_train.clone(), est.counts.copy(), s, lp, lp2)
    a, b = out["ref"], out["mirror"]
    assert torch.equal(a[0], b[0]) and (a[1] == b[1]).all()
    assert a[2].shape == (9, 2) and torch.equal(a[2], b[2]) and torch.allclose(a[3], b[3], atol=1e-6)
    assert torch.allclose(a[4], b[4], atol=1e-6) and torch.isfinite(b[4]).all()
