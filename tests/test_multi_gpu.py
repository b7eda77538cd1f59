"""Sharded runs on real GPUs over NCCL (needs >= 2 devices; skipped on a single-GPU box).

Particles are sharded in contiguous blocks, Philox draws are keyed by the global particle id and the
FIXED resampling scan is exact integer arithmetic, so a W-GPU run must reproduce the 1-GPU run: same
beta schedule and sweep counts, same ancestors (hence identical particle multisets up to the rounding of
the all-reduced moments, which enter the proposals at the 1e-13 level)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, model, out):
    import torch.distributed as dist
    import smcb200 as pkg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lik, prior = _problem(pkg, model)
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n), comm=pkg.TorchComm())
    eng.sample_prior()
    res = eng.run()
    np.savez(os.path.join(out, f"rank{rank}.npz"), particles=res.particles, lk=res.lk, betas=np.array(res.betas),
             n_mh=np.array(res.n_mh), moved=np.array(res.n_moved), logz=res.log_evidence, ess=np.array(res.ess),
             filled=np.array([s.filled for s in res.stages]))
    eng.close()
    dist.destroy_process_group()


def _problem(pkg, model):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if model == "mm_progress":
        g = np.load(os.path.join(root, "tests", "golden", "mm_reference_run.npz"))
        return pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"]), pkg.UniformBox([0, 0, 0], [10, 10, 10])
    return pkg.MMRate.synthetic(500), pkg.UniformBox([0, 0, 0], [10, 10, 10])


@pytest.mark.parametrize("model,n", [("mm_rate", 1 << 14), ("mm_progress", 1 << 14)])
def test_sharded_run_reproduces_single_gpu_run(pkg, tmp_path, model, n):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n, model, str(tmp_path)), nprocs=world, join=True)
    lik, prior = _problem(pkg, model)
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n))
    eng.sample_prior()
    ref = eng.run()
    eng.close()
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    P = np.concatenate([p["particles"] for p in parts])
    L = np.concatenate([p["lk"] for p in parts])
    for p in parts:                                        # every rank holds the same scalars
        assert np.array_equal(p["betas"], np.array(ref.betas))
        assert np.array_equal(p["n_mh"], np.array(ref.n_mh)) and np.array_equal(p["moved"], np.array(ref.n_moved))
        assert abs(float(p["logz"]) / ref.log_evidence - 1) < 1e-10
        assert np.abs(p["ess"] / np.array(ref.ess) - 1).max() < 1e-10
    assert P.shape == ref.particles.shape
    assert np.abs(P - ref.particles).max() < 1e-8 and np.abs(L / ref.lk - 1).max() < 1e-8
