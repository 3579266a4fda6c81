"""Two-GPU tests of the public multi-GPU API (`-m gpu`; skipped on boxes with fewer than two devices): one process per
GPU over NCCL, spawned from the test.  What the ranks check:

  (i)   after `prefill_sharded` (per-dimension prefills split over the ranks, slots broadcast over NCCL) every rank's logits
        are BIT-equal to those after a local `prefill` of all dimensions;
  (ii)  `distributed.sample_sharded`: the gathered result is the concatenation, in rank order, of what each rank draws on its
        own with its Philox row offset (so shards differ from each other and each equals a single-GPU run);
  (iii) `log_prob_sharded` equals the local `log_prob`; `sample_batched_sharded` equals a local `sample_batched` of the
        rank's block of observations.
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _rank_main(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.setdefault("NPE_PFN_B200_ALLOW_RANDOM_INIT", "1")
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        from npe_pfn_b200 import NPE_PFN_Core
        from npe_pfn_b200.distributed import log_prob_sharded, sample_batched_sharded, sample_sharded, shard_bounds
        from npe_pfn_b200.engine import Engine
        from npe_pfn_b200.weights import PFNWeights
        g = torch.Generator().manual_seed(77)
        N, dx, dth = 300, 3, 5
        theta = torch.randn(N, dth, generator=g)
        x = theta[:, :dx] * 0.7 + 0.2 * torch.randn(N, dx, generator=g)
        xo = x[:1].clone()
        Xt = torch.randn(90, dx + dth, generator=g)
        eng = Engine(weights=PFNWeights.random_init(), device=rank, max_slots=16)
        prior = torch.distributions.MultivariateNormal(torch.zeros(dth), 4.0 * torch.eye(dth))
        post = NPE_PFN_Core(prior=prior, regressor_init_kwargs={"engine": eng, "n_estimators": 1}).append_simulations(theta, x)
        # (i) sharded prefill vs local prefill
        post.prefill_sharded(xo)
        sharded = [eng.forward_logits(d, Xt[:, :dx + d]).clone() for d in range(dth)]
        post.invalidate_cache()
        post.prefill(xo)
        local = [eng.forward_logits(d, Xt[:, :dx + d]) for d in range(dth)]
        ok_prefill = all(torch.equal(a, b) for a, b in zip(sharded, local))
        # (ii) sample_sharded = rank-ordered concatenation of single-GPU runs with the rank's row offset
        total = 2 * 150 + 1
        post.shard_prefill = True
        torch.manual_seed(5)
        (s, lp), rate = sample_sharded(post, total, xo, with_log_prob=True, max_sampling_batch_size=64)
        lo, hi = shard_bounds(total, world, rank)
        torch.manual_seed(5)
        post.rank_row_offset = rank << 40
        own, own_lp = post.sample((hi - lo,), xo, with_log_prob=True, max_sampling_batch_size=64)
        ok_sample = s.shape == (total, dth) and s.device.type == "cpu" and torch.equal(s[lo:hi], own) and torch.equal(lp[lo:hi], own_lp)
        olo, ohi = shard_bounds(total, world, 1 - rank)
        ok_sample &= not torch.equal(s[lo:lo + 100], s[olo:olo + 100]) and abs(rate - 1.0) < 1e-6
        (sd, _), _ = sample_sharded(post, total, xo, with_log_prob=True, device_result=True)
        ok_sample &= sd.is_cuda and sd.shape == (total, dth)
        # (iii) log_prob_sharded / sample_batched_sharded
        th = torch.randn(41, dth, generator=g)
        ok_lp = torch.equal(log_prob_sharded(post, th, xo), post.log_prob(th, xo))
        xs = x[:5].clone()
        torch.manual_seed(9)
        sb = sample_batched_sharded(post, xs, 30)
        blo, bhi = shard_bounds(5, world, rank)
        torch.manual_seed(9)
        post.rank_row_offset = rank << 40
        own_b = post.sample_batched(xs[blo:bhi], (30,))
        ok_batched = sb.shape == (5, 30, dth) and torch.equal(sb[blo:bhi], own_b)
        dist.barrier()
        q.put((rank, bool(ok_prefill), bool(ok_sample), bool(ok_lp), bool(ok_batched), ""))
        dist.destroy_process_group()
    except Exception as e:  # report instead of hanging the parent
        import traceback
        q.put((rank, False, False, False, False, traceback.format_exc()[-1500:]))


def test_two_gpu_public_api():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for r in res:
        assert r[5] == "", r[5]
    assert [r[:5] for r in res] == [(0, True, True, True, True), (1, True, True, True, True)], res
