"""bench.py contract pieces that can be checked without a GPU: the reference arm (`--impl reference`, the oracle port on the host
cores) prints one JSON line with the agreed keys, non-zero ranks print nothing, and the GPU arm refuses to run without CUDA
instead of falling back to the CPU."""
import importlib
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bench(monkeypatch):
    b = importlib.import_module("bench")
    monkeypatch.setenv("DFGPU_Q3_SF", "0.2")          # a small instance of the same Q3 workload keeps the CPU arm to seconds here
    return b


def test_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    b = _bench(monkeypatch)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="reference")
    b.run_reference(args, 0, 1)
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == b.METRIC
    assert d["config"]["workload"].startswith("C4 TPC-H Q3-shaped pipeline, SF0.2") and d["config"]["fingerprint"][0] > 1000
    assert d["cpu_baseline"]["runs"] == 2 and d["cpu_baseline"]["best_rows_per_s"] >= d["cpu_baseline"]["median_rows_per_s"] * 0.999


def test_usable_threads_respects_affinity_and_quota(monkeypatch):
    b = importlib.import_module("bench")
    n = b.usable_threads()
    assert 1 <= n <= len(os.sched_getaffinity(0))
    assert b.cpu_sample_sf(100.0) in (100.0, 50.0, 25.0, 12.5, 6.25, 3.125, 1.5625, 0.78125)


def test_reference_arm_is_rank0_only(monkeypatch, capsys):
    b = _bench(monkeypatch)
    args = types.SimpleNamespace(gpus=2, steps=1, warmup=0, impl="reference")
    b.run_reference(args, 1, 2)
    assert capsys.readouterr().out.strip() == ""


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="needs a machine WITHOUT a GPU")
def test_gpu_arm_fails_loudly_without_cuda():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout) or "CUDA" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") and '"value"' in l for l in r.stdout.splitlines())
