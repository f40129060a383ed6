"""Host logic of bench.py that needs no GPU: the wall-budget plan and the command line."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_plan_steps_treats_steps_and_warmup_as_caps_under_the_budget():
    b = _bench()
    # the driver's call at 36 s per step: 3 warm-ups in total, as many timed steps as fit, never fewer than 3
    w_extra, k = b.plan_steps(20, 5, 36.0, 425.0)
    assert w_extra == 2 and 3 <= k <= 20 and (w_extra + k) * 36.0 <= 425.0
    # fast steps (8 GPUs): the requested counts are honoured
    assert b.plan_steps(20, 5, 4.5, 500.0) == (4, 20)
    # no budget left: still min(3, steps) timed steps, no extra warm-up
    assert b.plan_steps(20, 5, 36.0, 0.0) == (0, 3)
    assert b.plan_steps(1, 3, 36.0, 0.0) == (0, 1)
    # a single requested step keeps its warm-ups when they fit
    assert b.plan_steps(1, 3, 36.0, 400.0) == (2, 1)


def test_bench_without_cuda_fails_loudly():
    """No CPU fallback: the B200 arm refuses to run without a device."""
    import torch

    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "CUDA" in (res.stderr + res.stdout)
