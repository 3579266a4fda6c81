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
