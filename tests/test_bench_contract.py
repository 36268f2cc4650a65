"""bench.py's output contract, checked on the CPU through the reference arm (the only arm that runs without a GPU):
exactly ONE JSON line on stdout -- whatever libraries write to file descriptor 1 -- carrying the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


@pytest.mark.parametrize("workload", ["batch", "clip"])
def test_reference_arm_prints_one_json_line(workload):
    r = _run("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s"
    # "reference": the unmodified reference module from baseline/_ref (tools/prep_ref.py vendors it where /root/reference
    # exists; it travels to the GPU box with the snapshot); "port": the bit-identical oracle port when it is absent
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "Module2", "models", "networks.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_on_other_ranks_exits_quietly():
    r = _run("--impl", "reference", "--gpus", "2", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout == ""


def test_stdout_is_claimed_before_libraries_can_write_to_it():
    code = ("import sys, ctypes; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "ctypes.CDLL(None).puts(b'NCCL version 0.0'); print('stray'); bench.emit({'ok': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.stdout == '{"ok": 1}\n' and "NCCL version" in r.stderr and "stray" in r.stderr


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
