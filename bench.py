#!/usr/bin/env python
"""bench.py — posterior samples/sec (autoregressive, 10k simulations) on N B200s.

One "step" = one pass of the hot path on the `gaussian_linear` workload (BASELINE.json configs[1]): 10-D theta /
10-D x, N = 10 000 simulations as context, S posterior draws for one observation.  Every step rebuilds the
per-dimension K/V caches (10 prefills) and then runs the 10 autoregressive dimensions, so nothing is carried over
from a previous step.  With N GPUs every rank draws S samples (weak scaling, rows sharded, context replicated) and
the finished draws are all-gathered over NCCL inside the timed region; the per-dimension context prefills are split
over the ranks and the finished K/V-cache slots broadcast over NCCL (every rank still ends up with all ten caches).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl ours|reference]

`value`      device-timed (CUDA events, max over ranks) samples/s with the context already resident in HBM.
`e2e`        the same metric through the public API with HOST tensors: `append_simulations(theta, x)` (pinned H2D),
             `posterior.sample((S,), x_o)` and the draws copied back to the host, all inside the timed region.
`roofline`   item attention of test rows against the cached K/V (87 % of the FLOPs): algorithmic FLOPs / launch
             duration measured with CUDA events around each launch on its stream, against the measured bf16 peak.
`cpu_baseline` the CPU oracle (port of the reference's loop + tabpfn restatement) on a bounded sample.
`--impl reference` times that CPU path alone (the reference's arithmetic dependency `tabpfn` is not installable
offline, so the oracle port stands in; DESIGN.md "reference arm").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "posterior samples/sec (autoregressive, 10k simulations)"
DIM_X, DIM_THETA, N_CTX = 10, 10, 10_000
E, L, HID, BUCKETS = 192, 12, 768, 5000


# ---- workload ---------------------------------------------------------------------------------------------------
def make_workload(seed=42):
    """gaussian_linear (sbibm shape): theta ~ N(0, 0.1 I), x = theta + N(0, 0.1 I)."""
    g = torch.Generator().manual_seed(seed)
    theta = math_sqrt(0.1) * torch.randn(N_CTX, DIM_THETA, generator=g)
    x = theta + math_sqrt(0.1) * torch.randn(N_CTX, DIM_X, generator=g)
    theta_o = math_sqrt(0.1) * torch.randn(1, DIM_THETA, generator=g)
    x_o = theta_o + math_sqrt(0.1) * torch.randn(1, DIM_X, generator=g)
    prior = torch.distributions.MultivariateNormal(torch.zeros(DIM_THETA), 0.1 * torch.eye(DIM_THETA))
    return theta, x, x_o, prior


def math_sqrt(v):
    return float(v) ** 0.5


def tokens(d):
    return (DIM_X + d + 1) // 2 + 1


def flops_per_row(T, N):
    """Algorithmic FLOPs of one test row x one dimension as EXECUTED: SURVEY.md §8d's per-layer terms for layers
    0..L-2; in the last layer only the y-token column goes through item attention, out-projections and the MLP
    (the decoder reads nothing else), the feature-attention QKV / scores still cover all T tokens."""
    full = T * (28 * E * E + 4 * T * E + 4 * N * E)
    last = T * (12 * E * E + 4 * T * E) + (4 * E * E) + (4 * E * E + 4 * N * E) + 16 * E * E
    return (L - 1) * full + last + 2 * E * HID + 2 * HID * BUCKETS


def flops_prefill(T, N):
    """Context rows: full layers 0..L-2; the last layer stops after its K/V projection (final states are unused)."""
    full = 32 * N * T * E * E + 4 * N * T * T * E + 4 * T * N * N * E
    last = N * T * (16 * E * E + 12 * E * E) + 4 * N * T * T * E
    return (L - 1) * full + last


def flops_per_step(S):
    """FLOPs actually executed per step: 10 prefills + S rows x dims 1..9 + ONE row for dim 0 (all rows of
    dimension 0 are the same observation, so its logits are computed once)."""
    f = sum(flops_prefill(tokens(d), N_CTX) for d in range(DIM_THETA))
    f += flops_per_row(tokens(0), N_CTX)
    f += S * sum(flops_per_row(tokens(d), N_CTX) for d in range(1, DIM_THETA))
    return f


# ---- clocks -----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---- CPU baseline (oracle port; also the --impl reference arm) -----------------------------------------------------
def cpu_reference_throughput(m_rows=256, repeats=1):
    """Times the CPU oracle on a bounded sample of the workload and scales it to samples/s.

    Sample: the reference's `_sample` loop (npe_pfn.py:111-169: fit + predict + criterion.sample) restricted to
    parameter dimension 0 (6 of the 85 token columns) with `m_rows` draws at the full 10k-row context; the time of
    a full 10-dimension call with the reference's default max_sampling_batch_size = 10 000 is extrapolated with the
    FLOP model of SURVEY.md §8d (context re-fitted per dimension per call, as the reference does)."""
    from npe_pfn_b200.weights import PFNWeights
    from oracle.estimator import OracleTabPFNRegressor
    torch.set_num_threads(os.cpu_count() or 1)
    theta, x, x_o, _ = make_workload()
    w = PFNWeights.random_init()
    model = OracleTabPFNRegressor(weights=w, chunk=256)
    joint = torch.cat([x, theta], 1)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        model.fit(joint[:, :DIM_X], joint[:, DIM_X])
        pd = model.predict(x_o.repeat(m_rows, 1), output_type="full", quantiles=[])
        pd["criterion"].sample(pd["logits"])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    f_meas = flops_prefill(tokens(0), N_CTX) + m_rows * flops_per_row(tokens(0), N_CTX)
    batch = 10_000  # reference default max_sampling_batch_size: one fit per dimension per 10k draws
    f_call = sum(flops_prefill(tokens(d), N_CTX) + batch * flops_per_row(tokens(d), N_CTX) for d in range(DIM_THETA))
    t_call = best * f_call / f_meas
    return {"value": batch / t_call, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle fit+predict+sample of dimension 0 (T=6) with {m_rows} draws at N=10k context "
                      f"({best:.1f} s measured, {f_meas / best / 1e9:.0f} GFLOP/s), extrapolated by the FLOP model to a "
                      f"10-dimension call of 10k draws with a re-fit per dimension"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_throughput(m_rows=128)
        if i >= args.warmup:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    last = vals[-1]
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * 10_000 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gaussian_linear: 10-D theta / 10-D x, 10k simulations (bounded CPU sample)",
                   "n_estimators": 1},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- main arm ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=100_000, help="posterior draws per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attn", default=None, choices=[None, "mma", "tc"])
    ap.add_argument("--gemm", default=None, choices=[None, "mma", "tc"])
    ap.add_argument("--attn-poly", type=int, default=None)
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (tuning sweeps)")
    ap.add_argument("--item-gain", type=float, default=1.0,
                    help="tuning only: scale the item-attention Q/K projection weights (sharper attention scores)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    from npe_pfn_b200 import NPE_PFN_Core
    from npe_pfn_b200.engine import Engine
    from npe_pfn_b200.weights import PFNWeights

    S = args.samples
    theta, x, x_o, prior = make_workload()
    theta_p, x_p, xo_p = theta.pin_memory(), x.pin_memory(), x_o.pin_memory()
    w_bench = PFNWeights.random_init()
    if args.item_gain != 1.0:  # not the bench configuration: scores grow by gain^2 (reference-maximum changes per row)
        w_bench.t["item_wqkv"][:, :2 * w_bench.cfg.emsize] *= args.item_gain
    eng = Engine(weights=w_bench, device=local_rank, max_slots=16)
    if args.attn:
        eng.set_option("attn_impl", 1 if args.attn == "tc" else 0)
    if args.gemm:
        eng.set_option("gemm_impl", 1 if args.gemm == "tc" else 0)
    if args.attn_poly is not None:
        eng.set_option("attn_poly", args.attn_poly)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    post = NPE_PFN_Core(prior=prior, regressor_init_kwargs={"engine": eng})
    post.rank_row_offset = rank << 40
    post.shard_prefill = world > 1  # all ranks hold the same simulations and call sample() together
    post.append_simulations(theta_p, x_p)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """context resident in HBM; caches rebuilt; S draws; gather of the finished draws"""
        eng.__dict__.pop("_slot_tags", None)
        s, _ = post._sample(S, xo_p, return_device=True)
        if world > 1:
            s = _all_gather(s)
        return s

    def _all_gather(s):
        s = s.contiguous()  # the draws are a column slice of the joint test matrix
        out = torch.empty((world * s.shape[0],) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
        dist.all_gather_into_tensor(out, s)
        return out

    def step_e2e():
        """public API with host tensors: H2D of the simulations, sample(), D2H of the draws"""
        post.append_simulations(theta_p, x_p)
        out = post.sample((S,), xo_p, max_sampling_batch_size=S)
        return out

    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- timed region: device-resident ---------------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    clk = clocks.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = S * world * args.steps / (ms / 1e3)

    # ---- end to end through the public API (host buffers) -----------------------------------------------------------
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = S * world * args.steps / float(t.item())
    h2d = (theta.numel() + x.numel() + x_o.numel()) * 4
    d2h = S * DIM_THETA * 4

    # ---- roofline of the dominant kernel, measured live with CUDA events around each launch ----------------------
    eng.set_option("time_kernels", 1)
    eng.kernel_times(reset=True)
    step_device()
    kt = eng.kernel_times(reset=True)
    eng.set_option("time_kernels", 0)
    # ---- log_prob rows/s on the same workload (config 2: "100k samples plus autoregressive log_prob") --------------
    th_dev = step_device()[:S].contiguous()
    lp_rows = min(S, 50_000)
    post._autoregressive_log_prob(th_dev[:lp_rows], xo_p, return_device=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    post._autoregressive_log_prob(th_dev[:lp_rows], xo_p, return_device=True)
    e3.record()
    barrier()
    lp_ms = e2.elapsed_time(e3)
    logprob_rows_per_s = lp_rows * world / (lp_ms / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    a_ms, a_cnt, a_fl = kt["attn_test"]
    achieved = a_fl / (a_ms * 1e-3) / 1e12 if a_ms > 0 else 0.0
    tot_ms = sum(v[0] for v in kt.values())
    roofline = {"bound": "tensor", "kernel": "item attention of test rows vs cached K/V", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # DRAM bytes per launch from ncu (profiles/r1_launch_summary_v5_fused.csv: 167 attn_tc launches of 16 384
                # query rows read 12 393 MB and wrote 2 048 MB = 86.5 MB per launch; Q in + O out are 2 x 384 B per token,
                # the K/V tiles are L2 hits), scaled to this run's rows per launch (the library's chunk of 37 888 rows)
                "traffic": 8.65e7 * min(S, 37888) / 16384.0,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                               if peaks else "fallback 1.4 PFLOP/s sustained",
                "launches": a_cnt, "avg_launch_ms": a_ms / max(a_cnt, 1),
                "share_of_timed_kernels": a_ms / tot_ms if tot_ms else None,
                "per_class_ms": {k: round(v[0], 3) for k, v in kt.items()},
                "per_class_tflops": {k: (v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0) for k, v in kt.items()},
                "step_tflops": flops_per_step(S) / (ms_per_step * 1e-3) / 1e12,
                "step_frac_of_peak": flops_per_step(S) / (ms_per_step * 1e-3) / 1e12 / peak,
                # what actually bounds this kernel: one exponential per 128 tensor FLOPs (head dim 32).  MUFU alone
                # retires 15.9 ex2 / clk / SM (profiles/r1_pipe_rates_microbench.txt); the kernel splits the
                # exponentials between MUFU and the FMA pipes
                # the fused MLP sub-layer (mlp_tc.cuh) is the step's dense-GEMM kernel proper: 4 * tokens * 192 * 768 FLOPs per launch
                "mlp_kernel": ({"achieved": kt["mlp"][2] / (kt["mlp"][0] * 1e-3) / 1e12, "unit": "TFLOP/s",
                                "frac": kt["mlp"][2] / (kt["mlp"][0] * 1e-3) / 1e12 / peak, "launches": kt["mlp"][1]}
                               if kt.get("mlp", (0, 0, 0))[0] > 0 else None),
                "exponentials": {"achieved_per_s": achieved * 1e12 / 128.0,
                                 "mufu_only_peak_per_s": 15.9 * 148 * 1.965e9,
                                 "frac_of_mufu_only_peak": achieved * 1e12 / 128.0 / (15.9 * 148 * 1.965e9)}}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cpu = cpu_reference_throughput(m_rows=128)
            except Exception as ex:  # the baseline is a reported extra, never a reason to lose the GPU line
                cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                       "sample": f"failed: {ex}"}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "gaussian_linear: 10-D theta / 10-D x, 10k simulations, "
                                   f"{S} posterior draws per GPU per step via the autoregressive sampler",
                       "samples_per_gpu": S, "context_rows": N_CTX, "n_estimators": 1,
                       "weights": "seeded random init of the TabPFNv2 regressor architecture",
                       "prefill_in_step": True,
                       "l2": "per-step working set (1.3 GB of K/V cache + activations) exceeds the 126 MB L2",
                       "parallelism": f"rows sharded over {world} GPU(s), context replicated"
                                      + ("; per-dimension prefills split over the ranks, slots broadcast over NCCL"
                                         if world > 1 else "")},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "logprob": {"value": logprob_rows_per_s, "unit": "rows/s", "rows": lp_rows,
                        "note": "autoregressive log_prob of posterior draws, K/V caches reused, device resident"},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
