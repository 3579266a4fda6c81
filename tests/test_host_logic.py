"""CPU tests (`-m "not gpu"`) of the host side: the C-ABI library loads and exports every symbol the header
declares, the rejection loop / filters / helpers behave like the reference's, the product path refuses to run
without a GPU, and the row-sharding logic works over a world_size-2 gloo group."""
import os
import re
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_abi_symbols_match_header():
    from npe_pfn_b200.engine import ABI_SYMBOLS, load_library
    hdr = open(os.path.join(ROOT, "include", "npe_pfn_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pfn_[a-z_0-9]+)\s*\(", hdr)) - {"pfn_ctx"})
    assert sorted(ABI_SYMBOLS) == declared
    L = load_library()
    for s in declared:
        assert hasattr(L, s), s
    assert L.pfn_abi_version() == 2


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from npe_pfn_b200 import NPE_PFN_Core
    from npe_pfn_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()
    # a posterior object can be constructed and filled anywhere (like the reference's, and so it can be unpickled on a
    # login node); anything that computes needs the engine and fails loudly without a CUDA device - no CPU path
    prior = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    post = NPE_PFN_Core(prior=prior, regressor_init_kwargs={"n_estimators": 1})
    post.append_simulations(torch.randn(8, 2), torch.randn(8, 3))
    with pytest.raises(RuntimeError):
        post.engine
    with pytest.raises(RuntimeError):
        post.sample((4,), torch.randn(1, 3))
    with pytest.raises(RuntimeError):
        post.log_prob(torch.randn(4, 2), torch.randn(1, 3))
    with pytest.raises(RuntimeError):
        post.sample_batched(torch.randn(2, 3), (4,))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "npe_pfn_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "ctypes.CDLL(\"oracle" not in src and "libpfn_oracle" not in src, f


class _FakeProposal:
    """Deterministic proposal: row i of the k-th call is (k, i); accepted when (k + i) % 3 != 0."""

    def __init__(self, with_lp):
        self.calls = []
        self.with_lp = with_lp

    def __call__(self, n, **kw):
        k = len(self.calls)
        self.calls.append(n)
        c = torch.stack([torch.full((n,), float(k)), torch.arange(n, dtype=torch.float32)], 1)
        return c, (c.sum(1) if self.with_lp else None)


def _accept(c):
    return (c.sum(1).long() % 3) != 0


@pytest.mark.parametrize("num,max_bs,with_lp", [(10, 10_000, False), (1000, 300, True), (1, 5, False), (250, 100, True)])
def test_accept_reject_semantics(num, max_bs, with_lp):
    from npe_pfn_b200.accept_reject_sampler import accept_reject_sample
    p = _FakeProposal(with_lp)
    s, lp, rate = accept_reject_sample(p, _accept, num, max_sampling_batch_size=max_bs)
    assert s.shape == (num, 2) and bool(_accept(s).all())
    assert (lp is None) == (not with_lp)
    if with_lp:
        assert torch.equal(lp, s.sum(1))
    assert p.calls[0] == min(num, max_bs) and all(c <= max_bs for c in p.calls)
    assert 0 < rate <= 1
    if os.path.isdir(REF):  # same call trace and output as the unmodified reference loop
        sys.path.insert(0, REF)
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ref_ars", os.path.join(REF, "npe_pfn", "accept_reject_sampler.py"))
            ref = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref)
        finally:
            sys.path.remove(REF)
        p2 = _FakeProposal(with_lp)
        s2, lp2, rate2 = ref.accept_reject_sample(p2, _accept, num, max_sampling_batch_size=max_bs)
        assert p.calls == p2.calls and torch.equal(s, s2) and rate == rate2
        if with_lp:
            assert torch.equal(lp, lp2)


def test_accept_reject_max_iter_appends_unfiltered():
    from npe_pfn_b200.accept_reject_sampler import accept_reject_sample
    p = _FakeProposal(True)
    s, lp, _ = accept_reject_sample(p, lambda c: torch.zeros(len(c), dtype=torch.bool), 7, max_sampling_batch_size=50,
                                    max_iter_rejection=2)
    assert len(p.calls) == 3 and s.shape == (7, 2) and lp.shape == (7,)


def test_filters_and_helpers():
    from npe_pfn_b200.support_posterior import (check_for_uniform, get_filtering_method, get_uniform_bounds,
                                                prereject_with_bounds)
    from npe_pfn_b200.utils import BoxUniform, box_bounds_of, simulate_for_sbi
    g = torch.Generator().manual_seed(0)
    theta, x = torch.randn(500, 2, generator=g), torch.randn(500, 3, generator=g)
    obs = x[:1]
    f = get_filtering_method("standardized_euclidean_filtering")
    th, xx = f(obs, theta, x, 50)
    assert th.shape == (50, 2) and torch.equal(xx[0], x[0])  # nearest row is the observation itself
    z = (x - x.mean(0)) / x.std(0)
    d = torch.norm(z - (obs - x.mean(0)) / x.std(0), dim=1)
    assert torch.equal(xx, x[torch.topk(d, 50, largest=False)[1]])
    assert get_filtering_method("latest_filtering")(obs, theta, x, 20)[0].shape == (20, 2)
    assert get_filtering_method("no_filtering")(obs, theta, x, 20)[0].shape == (500, 2)
    assert get_filtering_method("random_filtering")(obs, theta, x, 20)[1].shape == (20, 3)
    with pytest.raises(ValueError):
        get_filtering_method("bogus")
    box = BoxUniform(-torch.ones(2), 2 * torch.ones(2))
    lo, hi = box_bounds_of(box)
    assert torch.equal(lo, -torch.ones(2)) and torch.equal(hi, 2 * torch.ones(2))
    assert box_bounds_of(torch.distributions.Uniform(-torch.ones(2), torch.ones(2)))[1].tolist() == [1, 1]
    assert box_bounds_of(torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))) == (None, None)
    assert box_bounds_of(torch.distributions.Normal(torch.zeros(2), torch.ones(2))) == (None, None)
    assert box_bounds_of(torch.distributions.Gamma(torch.ones(2), torch.ones(2))) is None
    assert check_for_uniform(box) and get_uniform_bounds(box)[0].tolist() == [-1, -1]
    c, rate = prereject_with_bounds(box, torch.zeros(2), torch.ones(2), 100, pre_sampling_batch_size=10_000)
    assert c.shape == (100, 2) and bool(((c >= 0) & (c <= 1)).all()) and 0.05 < rate < 0.2
    mvn = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    c, rate = prereject_with_bounds(mvn, -torch.ones(2), torch.ones(2), 100, pre_sampling_batch_size=1000)
    assert c.shape == (100, 2) and bool((c.abs() <= 1).all())
    th, xs = simulate_for_sbi(lambda t: t * 2, box, 10, simulation_batch_size=3)
    assert th.shape == (10, 2) and torch.equal(xs, th * 2)


def test_shard_bounds():
    from npe_pfn_b200.distributed import shard_bounds
    for total in (0, 1, 7, 8, 100_003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(total, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


class _FakePosterior:
    """Stands in for the GPU posterior: rank-dependent deterministic 'draws' so the gather order is checkable."""

    def __init__(self):
        self.rank_row_offset = 0
        self.last_acceptance_rate = 0.5

    def sample(self, shape, x, with_log_prob=False):
        n = shape[0]
        r = self.rank_row_offset >> 40
        s = torch.stack([torch.full((n,), float(r)), torch.arange(n, dtype=torch.float32)], 1)
        return (s, s.sum(1)) if with_log_prob else s

    def log_prob(self, theta, x):
        return theta.sum(1)

    _theta_train = torch.zeros(1, 2)

    def sample_batched(self, x, shape, with_log_prob=False):
        """observation o of this rank's block -> draws (value of x[o, 0], draw index)"""
        n = shape[0]
        s = torch.stack([x[:, :1].expand(-1, n), torch.arange(n, dtype=torch.float32).expand(x.shape[0], n)], -1)
        return (s, s.sum(-1)) if with_log_prob else s


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from npe_pfn_b200.distributed import (gather_rows, log_prob_sharded, reduce_counts, sample_batched_sharded, sample_sharded,
                                          shard_bounds)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        total = 11
        (s, lp), rate = sample_sharded(_FakePosterior(), total, torch.zeros(1, 2), with_log_prob=True)
        ok = s.shape == (total, 2) and lp.shape == (total,)
        for r in range(world):
            lo, hi = shard_bounds(total, world, r)
            ok &= bool((s[lo:hi, 0] == r).all()) and torch.equal(s[lo:hi, 1], torch.arange(hi - lo, dtype=torch.float32))
        ok &= abs(rate - 0.5) < 0.05
        th = torch.arange(14, dtype=torch.float32).reshape(7, 2)
        ok &= torch.equal(log_prob_sharded(_FakePosterior(), th, None), th.sum(1))
        ok &= reduce_counts(rank + 1, 10) == (sum(range(1, world + 1)), 10 * world)
        # observations split over the ranks, gathered back in observation order (also with fewer observations than ranks)
        for num_obs in (5, 1):
            xs = torch.arange(num_obs, dtype=torch.float32)[:, None].repeat(1, 3) + 100.0
            sb, sblp = sample_batched_sharded(_FakePosterior(), xs, 4, with_log_prob=True)
            ok &= sb.shape == (num_obs, 4, 2) and sblp.shape == (num_obs, 4)
            ok &= torch.equal(sb[:, :, 0], xs[:, :1].expand(-1, 4)) and torch.equal(sblp, sb.sum(-1))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_gather_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
