"""CPU tests (`-m "not gpu"`) of the ensemble path's host arithmetic (npe_pfn_b200/ensemble.py) against the
sklearn-backed oracle (oracle/ensemble.py): Yeo-Johnson maximum-likelihood lambdas vs sklearn's PowerTransformer,
the re-binning tables and their properties, member composition."""
import numpy as np
import pytest
import torch

from npe_pfn_b200 import ensemble as prod
from oracle import ensemble as orc


def _data(n=400, f=4, seed=0):
    g = np.random.default_rng(seed)
    X = g.normal(size=(n, f))
    X[:, 1] = np.exp(X[:, 1])            # skewed
    X[:, 2] = np.round(X[:, 2], 1)       # ties
    if f > 3:
        X[:, 3] = -np.abs(X[:, 3]) ** 1.5
    y = 0.5 * X[:, 0] + 0.3 * g.normal(size=n) + (X[:, 2] > 0) * 1.5
    return X, y


def test_members_same_as_oracle():
    for n in (2, 3, 8):
        a, b = prod.make_members(n, 7), orc.make_members(n, 7)
        assert [(m.x_kind, m.y_kind, m.perm_seed) for m in a] == [(m.x_kind, m.y_kind, m.perm_seed) for m in b]
        assert a[0].y_kind == "none"  # member 0 defines the common borders
    kinds = {(m.x_kind, m.y_kind) for m in prod.make_members(8)}
    assert len(kinds) == 4


def test_yeo_johnson_lambda_matches_sklearn():
    from sklearn.preprocessing import PowerTransformer, StandardScaler
    X, _ = _data()
    Z = StandardScaler().fit_transform(X)
    ref = PowerTransformer(method="yeo-johnson", standardize=False).fit(Z).lambdas_
    lam = prod.fit_yeo_johnson_lambda(torch.from_numpy(Z)).numpy()
    assert np.allclose(lam, ref, atol=2e-4), (lam, ref)
    # transform agrees with sklearn's at the fitted lambdas
    T = prod.yeo_johnson(torch.from_numpy(Z), torch.from_numpy(ref)).numpy()
    ref_T = PowerTransformer(method="yeo-johnson", standardize=False).fit(Z).transform(Z)
    assert np.allclose(T, ref_T, atol=1e-10)
    # constant column -> lambda 1
    Zc = np.concatenate([Z, np.full((Z.shape[0], 1), 3.0)], axis=1)
    assert prod.fit_yeo_johnson_lambda(torch.from_numpy(Zc))[-1].item() == 1.0


@pytest.mark.parametrize("lam", [-0.7, 0.0, 0.4, 1.0, 2.0, 2.6])
def test_yeo_johnson_inverse_round_trip(lam):
    x = torch.linspace(-4, 4, 101, dtype=torch.float64)
    y = prod.yeo_johnson(x, torch.tensor([lam], dtype=torch.float64))
    back = prod.yeo_johnson_inverse(y, lam)
    assert torch.allclose(back, x, atol=1e-9)
    assert np.allclose(orc.yeo_johnson(x.numpy(), lam), y.numpy(), atol=1e-12)
    inv_o = orc.yeo_johnson_inverse(np.linspace(-6, 6, 41), lam)
    inv_p = prod.yeo_johnson_inverse(torch.linspace(-6, 6, 41, dtype=torch.float64), lam).numpy()
    assert np.array_equal(np.isnan(inv_o), np.isnan(inv_p))
    assert np.allclose(inv_o[~np.isnan(inv_o)], inv_p[~np.isnan(inv_p)], atol=1e-9)


def test_rebin_tables_match_oracle_and_conserve_mass():
    from npe_pfn_b200.weights import default_borders
    z = default_borders(5000).double()
    for lam, shift in ((0.6, 0.1), (1.7, -0.2), (-0.3, 0.0), (2.5, 0.0)):
        bz = prod.yeo_johnson_inverse(z * 1.1 + shift, lam)
        idx, frac, valid = prod.rebin_tables(bz, z)
        oi, of, ov = orc.rebin_tables(bz.numpy(), z.numpy())
        assert np.array_equal(idx.numpy(), oi) and np.array_equal(valid.numpy().astype(bool), ov)
        assert np.allclose(frac.numpy(), of, atol=1e-6)
        g = np.random.default_rng(1)
        p = g.random((3, 5000))
        q = orc.translate_probs(p, oi, of, ov)
        assert q.min() >= 0 and q.sum(axis=1).max() <= 1 + 1e-9
        # mass that falls inside the common range is kept: compare with the member CDF at the two end borders
        pv = np.where(ov, p, 0.0)
        pv /= pv.sum(axis=1, keepdims=True)
        C = np.concatenate([np.zeros((3, 1)), np.cumsum(pv, axis=1)], axis=1)[:, :-1]
        inside = (C[:, oi[-1]] + pv[:, oi[-1]] * of[-1]) - (C[:, oi[0]] + pv[:, oi[0]] * of[0])
        assert np.allclose(q.sum(axis=1), inside, atol=1e-9)


def test_rebin_identity_borders_is_identity():
    from npe_pfn_b200.weights import default_borders
    z = default_borders(5000).double()
    idx, frac, valid = prod.rebin_tables(z, z)
    p = np.random.default_rng(2).random((2, 5000))
    p /= p.sum(axis=1, keepdims=True)
    q = orc.translate_probs(p, idx.numpy(), frac.numpy(), valid.numpy().astype(bool))
    assert np.allclose(q, p, atol=1e-12)
    out = orc.combine([np.log(p).astype(np.float32)] * 3, [None, (idx.numpy(), frac.numpy(), valid.numpy().astype(bool)), None])
    assert np.allclose(out, np.log(p), atol=1e-5)


def test_oracle_quantile_member_is_sklearn():
    """The oracle's quantile pipeline is sklearn's QuantileTransformer itself: uniform output in [0, 1], original
    columns appended, SVD components and fingerprint present, shuffle is a permutation of the columns."""
    X, _ = _data(200, 3)
    m = orc.OracleMember(orc.MemberSpec("quantile", "none", 5)).fit_x(X)
    T = m.transform_x(X)
    k = m.svd_k
    assert T.shape == (200, 2 * 3 + k + 1)
    inv = np.argsort(np.random.default_rng(5).permutation(T.shape[1]))
    base = T[:, inv]
    assert base[:, :3].min() >= 0 and base[:, :3].max() <= 1
    assert np.allclose(base[:, 3:6], X.astype(np.float32))
    assert np.allclose(base[:, -1], orc.fingerprint(X))
    fp = orc.fingerprint(np.array([[0.0, 1.0], [-0.0, 1.0], [1.0, 0.0]]))
    assert fp[0] == fp[1] != fp[2] and 0 <= fp.min() and fp.max() < 1


def test_ensemble_golden_vectors(weights):
    """Frozen outputs of the ensemble oracle (tests/golden/make_golden_ensemble.py): guards against drift of the
    oracle itself (sklearn / numpy versions on the box that runs the tests)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "ensemble_golden.pt")
    assert os.path.exists(path), "run tests/golden/make_golden_ensemble.py"
    gold = torch.load(path, weights_only=False)
    m = orc.OracleEnsembleRegressor(weights=weights, n_estimators=4, random_state=0).fit(gold["Xc"], gold["yc"])
    assert [(s.x_kind, s.y_kind, s.perm_seed) for s in m.specs] == gold["specs"]
    for mm, feat, lam in zip(m.members, gold["features"], gold["lam_y"]):
        assert np.allclose(mm.transform_x(gold["Xt"].numpy()), feat.numpy(), atol=2e-5)
        assert (lam is None) == (mm.lam_y is None) and (lam is None or abs(lam - mm.lam_y) < 1e-6)
    lg = m.predict(gold["Xt"])["logits"]
    assert torch.allclose(lg[:, gold["cols"]], gold["logits_cols"], atol=2e-3)
    assert torch.allclose(torch.logsumexp(lg, -1), gold["lse"], atol=1e-4)
