"""bench.py's reference arm (the CPU implementation of the path) runs without a GPU and prints exactly one JSON
line carrying the keys the bench contract names (CPU tests).  The main arm is run with the DRIVER's own argv on
the GPU box (`-m gpu`): round 1 shipped a bench that only ever ran with its default --steps 3 and died at
--steps 20 --warmup 5."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "particle_loglik_evals_per_s" and j["unit"] == "evals/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def _check_main_line(out, n_gpus, steps, warmup):
    assert out.returncode == 0, (out.stdout[-400:], out.stderr[-1500:])
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["metric"] == "particle_loglik_evals_per_s" and j["unit"] == "evals/s" and "impl" not in j
    assert j["n_gpus"] == n_gpus and j["steps"] == steps and j["warmup"] == warmup
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["higher_is_better"] is True and j["scaling"] == "weak"
    assert j["dtype"] == "f64" and "workload" in j["config"] and j["gpu_launches"] > 0
    r = j["roofline"]
    assert r["frac"] is not None and 0 < r["frac"] < 1 and r["achieved"] > 0 and r["peak"] > 0
    # the kernel time and the flop count behind `achieved` cover the same sweeps: every profiled sweep was read
    assert r["launches"] == j["likelihood_sweeps_timed"], (r["launches"], j["likelihood_sweeps_timed"])
    if n_gpus == 1:
        assert j["roofline_hbm"]["frac"] > 0
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = j["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and isinstance(c["reasons"], list)
    return j


@pytest.mark.gpu
def test_main_arm_survives_the_drivers_argv_on_one_gpu():
    """`python bench.py --gpus 1 --steps 20 --warmup 5`: ~20 x 34 likelihood sweeps under SMCB_PARAM_PROFILE."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "20", "--warmup",
                          "5"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    j = _check_main_line(out, 1, 20, 5)
    assert j["likelihood_sweeps_timed"] > 512      # more sweeps than the old fixed event ring held
    cb = j["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] == "port" and cb["sample"]


@pytest.mark.gpu
def test_main_arm_survives_the_drivers_argv_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--gpus", "2", "--steps", "20", "--warmup", "5"], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    _check_main_line(out, 2, 20, 5)
