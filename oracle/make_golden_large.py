"""Frozen golden vectors AT SIZE: the scaled-down configs 2 and 3 at N = 4,096 run through
the REAL reference, and config 2 at its full size N = 16,384 run through the oracle.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference for the N = 4,096 problems).

    python -m oracle.make_golden_large [c2_4096] [c3_4096] [c2_16384]

Only posterior quantities are stored (representer weights, mean / variance on the test grid, a 16 x 16 covariance
block): the Gram matrices themselves (134 MB / 2.1 GB) are covered entry-wise by ``kernels.npz``.  The problem is
NOT stored either -- it is regenerated from ``problem_spec`` by ``large_problem`` below (seeded), here and in the tests.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gp as ogp  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SPECS = {
    # config 2 scaled down: 3,840 PDE points + 4 x 64 boundary points, l = 4 / sqrt(N_pde)
    "c2_4096": {"kind": "poisson2d", "n_pde": 3840, "n_bc_edge": 64, "seed": 0, "grid": 32, "source": "reference"},
    # config 3 scaled down: heat equation, 64 IC + 2 x 96 BC + 96 x 40 PDE points
    "c3_4096": {"kind": "heat", "n_ic": 64, "n_bc": 96, "nt": 96, "nx": 40, "alpha": 0.1, "grid": 32, "source": "reference"},
    # config 2 at full size: 15,360 PDE points + 4 x 256 boundary points
    "c2_16384": {"kind": "poisson2d", "n_pde": 15360, "n_bc_edge": 256, "seed": 0, "grid": 64, "source": "oracle"},
}


def large_problem(spec):
    if spec["kind"] == "poisson2d":
        return ogp.poisson2d_problem(spec["n_pde"], spec["n_bc_edge"], seed=spec["seed"], grid=spec["grid"])
    return ogp.heat_problem(n_ic=spec["n_ic"], n_bc=spec["n_bc"], nt=spec["nt"], nx=spec["nx"], alpha=spec["alpha"],
                            grid=spec["grid"])


def main(names):
    for name in names:
        spec = SPECS[name]
        prob = large_problem(spec)
        n = sum(len(b["Y"]) for b in prob["blocks"])
        t0 = time.perf_counter()
        ores = ogp.solve(prob)
        t_or = time.perf_counter() - t0
        extra = {}
        if spec["source"] == "reference":
            from oracle import make_golden

            t0 = time.perf_counter()
            res = make_golden.run_reference_gp(prob)
            extra["reference_seconds"] = time.perf_counter() - t0
            for key in ("w", "mean", "var", "cov"):
                sc = max(np.max(np.abs(res[key])), np.max(np.abs(res["var"])))
                dev = np.max(np.abs(res[key] - ores[key])) / sc
                extra[f"oracle_vs_reference_{key}"] = dev
                print(f"{name:10s} N={n} {key:5s} oracle-ref rel {dev:.2e}")
        else:
            res = ores
        meta = dict(spec, N=n, oracle_seconds=t_or, **extra)
        print(name, json.dumps(meta))
        np.savez_compressed(
            os.path.join(GOLDEN, f"large_{name}.npz"),
            problem_spec=np.frombuffer(json.dumps(spec).encode(), dtype=np.uint8),
            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
            **{k: res[k] for k in ("w", "mean", "var", "cov")},
        )


if __name__ == "__main__":
    main(sys.argv[1:] or list(SPECS))
