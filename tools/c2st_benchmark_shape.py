#!/usr/bin/env python
"""One-off evidence run (too slow for the test suite): sample-set parity AT THE BENCHMARK SHAPE.

gaussian_linear of bench.py (10-D theta / 10-D x, 10 000 simulations): M draws of the CUDA sampler against M draws of the
oracle's restatement of the reference loop (all ten dimensions, re-fit per dimension, fp32 CPU, different uniforms), compared
with the reference's classifier two-sample test recipe (tests/c2st.py) and per-dimension KS tests.

    python tools/c2st_benchmark_shape.py [M] > profiles/r2_c2st_benchmark_shape.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NPE_PFN_B200_ALLOW_RANDOM_INIT", "1")

import torch  # noqa: E402

from bench import make_workload  # noqa: E402
from c2st import c2st, ks_pvalues  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    from npe_pfn_b200 import NPE_PFN_Core
    from npe_pfn_b200.engine import Engine
    from npe_pfn_b200.weights import PFNWeights
    from oracle.estimator import OracleTabPFNRegressor
    from oracle.reference_loop import sample_loop
    torch.set_num_threads(os.cpu_count() or 1)
    theta, x, x_o, prior = make_workload()
    w = PFNWeights.random_init()
    eng = Engine(weights=w, device=0)
    post = NPE_PFN_Core(prior=prior, regressor_init_kwargs={"engine": eng, "n_estimators": 1}).append_simulations(theta, x)
    torch.manual_seed(1)
    t0 = time.perf_counter()
    cuda = post.sample((M,), x_o)
    t_cuda = time.perf_counter() - t0
    cuda2 = post.sample((M,), x_o)  # a second, independent CUDA set: the C2ST noise floor
    torch.manual_seed(2)
    t0 = time.perf_counter()
    ref, _ = sample_loop(OracleTabPFNRegressor(weights=w, chunk=256), x, theta, x_o, M)
    t_ref = time.perf_counter() - t0
    print(f"# gaussian_linear, N = 10000 simulations, {M} draws each; CUDA {t_cuda:.2f} s, CPU oracle loop {t_ref:.1f} s")
    print(f"C2ST(CUDA, oracle loop)      = {c2st(ref, cuda, epochs=40, batch_size=256):.4f}")
    print(f"C2ST(CUDA, CUDA other seed)  = {c2st(cuda2, cuda, epochs=40, batch_size=256):.4f}   (noise floor of the recipe)")
    pv = ks_pvalues(ref, cuda)
    print("KS p-values per dimension (CUDA vs oracle loop):", [round(p, 3) for p in pv])
    print("mean |CUDA - oracle| of the per-dimension means / posterior std:",
          [round(float(v), 3) for v in ((cuda.mean(0) - ref.mean(0)).abs() / ref.std(0))])
    print("ratio of per-dimension stds (CUDA / oracle):", [round(float(v), 3) for v in (cuda.std(0) / ref.std(0))])


if __name__ == "__main__":
    main()
