"""bench.py contract checks that run without a GPU: the reference arm (the reference's own CPU code,
oracle/_ref) prints exactly one JSON line with the agreed keys, and the GPU arm refuses to run -- loudly --
when there is no CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(ref):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm runs the GPU arm's workload (whole 3840x2160 images, every host thread one) and says so
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    assert d["config"]["workload"] == bench.WORKLOAD_4K
    assert d["config"]["images_per_step"] == d["cpu_baseline"]["cores"]
    assert d["cpu_baseline"]["stock_single_thread"]["cores"] == 1 and d["cpu_baseline"]["stock_single_thread"]["value"] > 0
    assert "whole synthetic 3840x2160 image" in d["cpu_baseline"]["sample"]


def test_graph_length_divides_the_timed_steps():
    """No timed step may run eagerly: the CUDA graph holds K steps or a divisor of K (or 32 + a remainder graph)."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    for k in (1, 3, 20, 64, 100, 2000, 20000, 97, 101, 2003):
        L = bench._graph_len(k)
        assert 1 <= L <= 64
        assert k <= 64 and L == k or k % L == 0 or L == 32, (k, L)


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr or "no usable" in r.stderr
