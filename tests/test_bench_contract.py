"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the reference's own CPU
sampler on the host cores over a bounded sample of the bench workload) prints ONE JSON line carrying the keys the driver reads, with
the same metric / unit / workload string as the B200 arm would; without a CUDA device the B200 arm refuses instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fragment_assignment_draws_per_sec" and d["unit"] == "draws/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["ms_per_step"] > 0 and d["value"] > 0
    assert d["config"]["workload"].startswith("config 1 (BASELINE.json configs[0]): 200 loci x 1 chain(s)")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "loci of config 1" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the B200 arm runs (tested by the driver)")
    r = _run(["--config", "1", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-other-configs"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
