#!/usr/bin/env python
"""Where a sharded run spends its time outside the likelihood kernels (host clock with device syncs, rank 0).

    python profiles/multi_gpu_breakdown.py [world=2] [log2_particles_per_gpu=20]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, lg):
    import torch.distributed as dist
    import smcb200 as pkg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "29533"
    os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
    lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
    prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
    n = (1 << lg) * world
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n), comm=pkg.TorchComm())
    acc = {}

    def wrap(name):
        f = getattr(eng, name)

        def g_(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f(*a, **k)
            torch.cuda.synchronize()
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            acc[name + "_n"] = acc.get(name + "_n", 0) + 1
            return r
        setattr(eng, name, g_)

    for name in ("temper", "resample", "mh_sweep", "_launch_moments", "sim_particle"):
        wrap(name)
    for rep in range(2):
        acc.clear()
        eng.sample_prior()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = eng.run()
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
    if rank == 0:
        print(f"world={world} N={n}: total {total * 1e3:.1f} ms (device {res.seconds * 1e3:.1f} ms), {len(res.betas)} stages, {sum(res.n_mh)} sweeps")
        for k in ("sim_particle", "temper", "resample", "_launch_moments", "mh_sweep"):
            print(f"  {k:16s} {acc.get(k, 0) * 1e3:8.2f} ms over {acc.get(k + '_n', 0)} calls")
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(world, lg), nprocs=world, join=True)
