"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the
agreed keys (it times the compiled reference / the C port on the host cores), and the CUDA arm
refuses to run without a device instead of falling back to anything. CPU only."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--batch", "4", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines            # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "alexnet224_int8_images_per_s" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "alexnet224_int8_b4"


def test_reference_arm_other_ranks_do_nothing():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_cuda_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
