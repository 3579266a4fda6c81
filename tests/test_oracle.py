"""CPU tests of the oracle itself (`-m "not gpu"`): the C head against the textbook formulation and
Random123's Philox known-answer vectors, the split prefill/forward against the monolithic forward, and
the frozen golden vectors."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

from oracle import bar_head
from oracle import tabpfn_oracle as model
from oracle.estimator import OracleTabPFNRegressor, y_standardise
from oracle.reference_loop import logprob_loop, sample_loop

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    """Philox4x32-10 KAT (Random123 kat_vectors): counter / key -> output words."""
    L = bar_head.lib()
    out = (ctypes.c_uint32 * 4)()

    def run(ctr, key):
        seed = key[0] | (key[1] << 32)
        row = ctr[0] | (ctr[1] << 32)
        off = ctr[2] | (ctr[3] << 32)
        L.pfn_oracle_philox4x32(seed, row, off, out)
        return [int(v) for v in out]

    assert run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert run([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_exp_det_accuracy():
    L = bar_head.lib()
    ts = np.concatenate([np.linspace(-60, 0, 2001), -np.logspace(-6, 1.7, 200)]).astype(np.float32)
    got = np.array([L.pfn_oracle_exp_det(float(t)) for t in ts], dtype=np.float64)
    ref = np.exp(ts.astype(np.float64))
    rel = np.abs(got - ref) / ref
    # the argument reduction t*log2(e) is done in fp32, so the relative error grows with |t|
    assert np.max(rel) < 5e-6
    assert np.max(rel[ts > -4]) < 5e-7
    assert L.pfn_oracle_exp_det(0.0) == 1.0


@pytest.mark.parametrize("B", [8, 100, 5000])
def test_head_sample_matches_textbook(B):
    g = torch.Generator().manual_seed(B)
    M = 64
    logits = torch.randn(M, B, generator=g) * 3.0
    borders = torch.sort(torch.randn(B + 1, generator=g))[0]
    u = torch.rand(M, generator=g)
    theta, idx, _ = bar_head.sample(logits, borders, uniforms=u)
    t_ref, i_ref = bar_head.textbook_icdf(logits, borders, u)
    # identical bucket except where u falls within rounding distance of a CDF knot
    mism = (idx != i_ref).float().mean().item()
    assert mism <= 0.02
    ok = idx == i_ref
    width = (borders[1:] - borders[:-1])[idx.long()]
    assert torch.all((theta[ok] - t_ref[ok]).abs() <= 1e-3 * width[ok] + 1e-6)
    assert torch.all(theta >= borders[idx.long()]) and torch.all(theta <= borders[idx.long() + 1])


def test_head_sample_edge_cases():
    B = 16
    borders = torch.linspace(-1, 1, B + 1)
    # one-hot logits: always that bucket; uniform position inside equals u
    logits = torch.full((4, B), -1e4)
    logits[:, 5] = 0.0
    u = torch.tensor([1e-7, 0.25, 0.5, 0.999])
    theta, idx, _ = bar_head.sample(logits, borders, uniforms=u)
    assert idx.tolist() == [5, 5, 5, 5]
    lo, hi = borders[5].item(), borders[6].item()
    assert torch.allclose(theta, lo + (hi - lo) * u, atol=1e-6)
    # u == 0 (only reachable with injected uniforms): searchsorted-left lands on bucket 0, position 0
    theta0, idx0, _ = bar_head.sample(logits[:1], borders, uniforms=torch.zeros(1))
    assert int(idx0[0]) == 0 and float(theta0[0]) == float(borders[0])
    # -inf logits and flat logits
    logits = torch.zeros(2, B)
    logits[0, :8] = -float("inf")
    theta, idx, _ = bar_head.sample(logits, borders, uniforms=torch.tensor([1e-7, 0.5]))
    assert int(idx[0]) == 8 and int(idx[1]) == 7  # searchsorted-left: C_7 == target is not "below"
    # Philox uniforms are in [0, 1) and reproducible
    u1 = bar_head.philox_uniforms(7, 0, 3, 100)
    u2 = bar_head.philox_uniforms(7, 0, 3, 100)
    assert torch.equal(u1, u2) and float(u1.min()) > 0 and float(u1.max()) < 1


@pytest.mark.parametrize("B", [8, 5000])
def test_head_nll_matches_textbook(B):
    g = torch.Generator().manual_seed(100 + B)
    M = 200
    logits = torch.randn(M, B, generator=g) * 2.0
    borders = torch.sort(torch.randn(B + 1, generator=g))[0]
    lo, hi = borders[0].item(), borders[-1].item()
    y = torch.rand(M, generator=g) * (hi - lo) * 1.4 + lo - 0.2 * (hi - lo)  # includes both tails
    y[0], y[1] = borders[3], borders[-1]
    got = bar_head.nll(logits, borders, y)
    ref = bar_head.textbook_nll(logits, borders, y)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)


def test_nll_integrates_to_one():
    """exp(-nll) is a density: its integral over a fine grid (incl. half-normal tails) is ~1."""
    B = 32
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(1, B, generator=g)
    borders = torch.sort(torch.randn(B + 1, generator=g))[0]
    ys = torch.linspace(borders[0].item() - 6, borders[-1].item() + 6, 400001)
    dens = torch.exp(-bar_head.nll(logits.expand(ys.numel(), B).contiguous(), borders, ys).double())
    integral = torch.trapezoid(dens, ys.double()).item()
    assert abs(integral - 1.0) < 2e-3


@pytest.mark.parametrize("F", [1, 2, 3, 5])
def test_prefill_plus_forward_equals_joint(weights, F):
    g = torch.Generator().manual_seed(F)
    N, M = 24, 7
    Xc = torch.randn(N, F, generator=g)
    yc = torch.randn(N, generator=g)
    Xt = torch.randn(M, F, generator=g)
    cache = model.prefill(weights, Xc, yc, explicit=True)
    a = model.forward_test(weights, cache, Xt, explicit=True)
    b = model.forward_joint(weights, Xc, yc, Xt)
    assert a.shape == (M, weights.cfg.num_buckets)
    assert torch.allclose(a, b, atol=2e-4, rtol=1e-4)
    # test rows are independent of each other
    c = model.forward_test(weights, cache, Xt[2:3], explicit=True)
    assert torch.allclose(a[2:3], c, atol=1e-5)


def test_encoder_edge_cases(weights):
    """NaN / inf cells, a constant column and the zero padding column (odd F)."""
    g = torch.Generator().manual_seed(11)
    N, F = 20, 3
    Xc = torch.randn(N, F, generator=g)
    Xc[:, 1] = 2.5  # constant on the context
    Xc[3, 0] = float("nan")
    Xc[4, 2] = float("inf")
    yc = torch.randn(N, generator=g)
    st = model.EncoderStats(Xc, yc, 2)
    assert st.std[1] == 0 and st.std[3] == 0 and st.scale[0] == math.sqrt(2.0) and st.scale[1] == math.sqrt(2.0)
    Xt = torch.randn(5, F, generator=g)
    Xt[0, 0] = float("-inf")
    cache = model.prefill(weights, Xc, yc)
    out = model.forward_test(weights, cache, Xt)
    assert torch.isfinite(out).all()


def test_y_standardise_degenerate():
    m, s, yz = y_standardise(torch.full((5,), 3.0))
    assert m == 3.0 and s == 1.0 and torch.all(yz == 0)
    m, s, yz = y_standardise(torch.tensor([1.5]))
    assert s == 1.0


def test_reference_loops_small(weights):
    """sample_loop / logprob_loop (restating npe_pfn.py:111-169 / :462-524) are mutually consistent:
    the log-prob returned while sampling equals logprob_loop of the same draws."""
    g = torch.Generator().manual_seed(5)
    N, dx, dth, M = 16, 2, 2, 6
    theta = torch.randn(N, dth, generator=g)
    x = theta @ torch.randn(dth, dx, generator=g) + 0.1 * torch.randn(N, dx, generator=g) + 1.0
    xo = x[:1]
    u = torch.rand(M, dth, generator=g)
    m = OracleTabPFNRegressor(weights=weights)
    s, lp, bins = sample_loop(m, x, theta, xo, M, with_log_prob=True, uniforms=u, return_bins=True)
    assert s.shape == (M, dth) and bins.shape == (M, dth) and torch.isfinite(lp).all()
    lp2 = logprob_loop(m, x, theta, xo, s)
    assert torch.allclose(lp, lp2, atol=1e-4)


def test_golden_vectors(weights):
    """Frozen oracle outputs (tests/golden/make_golden.py): guards the oracle against drift."""
    path = os.path.join(GOLDEN, "oracle_golden.pt")
    assert os.path.exists(path), "run tests/golden/make_golden.py"
    gold = torch.load(path)
    for case in gold["cases"]:
        m = OracleTabPFNRegressor(weights=weights)
        m.fit(case["Xc"], case["yc"])
        logits = m.predict(case["Xt"])["logits"]
        assert torch.allclose(logits[:, case["cols"]], case["logits_cols"], atol=2e-4, rtol=1e-4)
        assert torch.allclose(torch.logsumexp(logits, -1), case["lse"], atol=2e-4)
        crit = m.predict(case["Xt"])["criterion"]
        theta, idx, _ = crit.icdf_indices(case["logits_full0"], case["u"][:1])
        assert int(idx[0]) == int(case["idx0"]) and float(theta[0]) == float(case["theta0"])
        nll = crit(logits, case["y"])
        assert torch.allclose(nll, case["nll"], atol=2e-4)


def test_fast_bucket_mass_arithmetic_equals_specification(tmp_path):
    """head_row2_kernel computes a bucket mass in 15 instructions (round-down magic add for floor, exponent insertion by an
    integer add, one F2I.U64.TRUNC).  tests/head_arith_check.c emulates exactly that sequence on the CPU and compares it with
    pfn_oracle_quantize(pfn_oracle_exp_det(t)) of oracle/bar_head.c: every 16th fp32 value of [-64, 0] here (70 M
    arguments, ~4 s) plus the neighbourhoods of every integer step, below -64, -inf, NaN; stride 1 (all 1.1e9) was run
    once by hand: 0 mismatches."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "head_arith_check")
    subprocess.check_call([gcc, "-O2", "-ffp-contract=off", "-frounding-math", os.path.join(here, "head_arith_check.c"),
                           os.path.join(here, "..", "oracle", "bar_head.c"), "-lm", "-o", exe])
    out = subprocess.run([exe, "16"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout and int(out.stdout.split()[1]) > 69_000_000


def test_integer_target_comparison_equals_specification():
    """head_row2_kernel replaces every `(double)C < target` of the specification (oracle/bar_head.c:116, target =
    (double)u * (double)Z) by the integer comparison `C < ceil(target)`.  Equivalent for every integer C below 2^53:
    checked here in exact arithmetic around the target for random partition sums Z and fp32 uniforms u, including u = 0,
    the largest u torch.rand can return, and targets that are integers themselves."""
    import random
    rng = random.Random(5)
    fr = np.float32
    us = [0.0, float(fr(2.0 ** -24)), float(fr(1.0 - 2.0 ** -24)), 0.5, 0.25]
    us += [float(fr(rng.random())) for _ in range(400)]
    for u in us:
        for _ in range(25):
            Z = rng.randrange(1 << 40, 5000 << 40)
            if rng.random() < 0.2:
                Z = (Z >> 30) << 30  # make u * Z an integer more often
            target = float(u) * float(Z)          # double product, as in the kernels (Z < 2^53 is exact as a double)
            tceil = math.ceil(target)              # exact: Python converts the double to an integer without rounding
            base = int(target)
            for C in {0, 1, Z, base - 2, base - 1, base, base + 1, base + 2, tceil - 1, tceil, tceil + 1}:
                if 0 <= C < (1 << 53):
                    assert (float(C) < target) == (C < tceil), (u, Z, C)
