"""Multi-rank host logic on the CPU (gloo, world_size 2 and 3): the engine's cross-shard resampling
(`engine.sharded_resample`: all-gather of shard totals -> integer migration plan -> all-to-all of contiguous
particle chunks) driven by NumPy twins of the device kernels must reproduce the unsharded oracle result.

The product never runs this backend: `_OracleOps` stands in for the CUDA kernels only so that the exchange
logic, which is the same code the GPUs run, can be exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(N, d, seed, conc):
    rs = np.random.RandomState(seed)
    w = rs.dirichlet(np.full(N, conc))
    state = rs.normal(0, 1, (d + 1, N))
    return w, state


class _OracleOps:
    """NumPy stand-ins for smcb_resample_totals / _counts / smcb_ancestors / smcb_gather on one shard."""

    def __init__(self, smc, w_shard, state_shard, N, u0, rank):
        self.smc, self.w, self.state, self.N, self.u0, self.rank = smc, w_shard, state_shard, N, u0, rank
        self.counts = None

    def totals(self):
        _, fl, q = self.smc.resample_fixed_shard(self.w, self.u0, self.N, 0, True)
        # q_total needs exact 64-bit storage: it is < 2^62 because the residuals of one shard sum to < 1
        return torch.tensor([fl, q], dtype=torch.int64)

    def counts_fixed(self, carry_q):
        self.counts, _, _ = self.smc.resample_fixed_shard(self.w, self.u0, self.N, carry_q, self.rank == 0)

    def counts_sequential(self, carry):
        self.counts, carry_out, fl, nc = self.smc.resample_sequential_shard(self.w, carry, self.N)
        return np.array(carry_out), torch.tensor([fl, nc], dtype=torch.int64)

    def expand_and_pack(self, m_loc, send, sendbuf):
        anc = self.smc.fit_ancestors(np.repeat(np.arange(len(self.w)), self.counts), m_loc)
        D1 = self.state.shape[0]
        off = 0
        for cnt in send:
            if cnt:
                chunk = self.state[:, anc[off:off + cnt]]                     # [D1][cnt], contiguous per destination
                sendbuf[D1 * off: D1 * (off + cnt)].copy_(torch.from_numpy(np.ascontiguousarray(chunk)).reshape(-1))
                off += cnt


class _OracleFusedOps(_OracleOps):
    """... plus the stand-in of the single-pass kernel: with it `sharded_resample` takes the device path (one kernel from
    weights to the rows of the shard's output slots, rows exchanged without packing)."""

    def fused_expand(self, carry_q, m_loc, sendbuf, ld_send):
        """Stand-in of smcb_resample_fused in its sharded form: rows of this shard's first m_loc output slots."""
        counts, _, _ = self.smc.resample_fixed_shard(self.w, self.u0, self.N, carry_q, self.rank == 0)
        anc = self.smc.fit_ancestors(np.repeat(np.arange(len(self.w)), counts), m_loc)
        D1 = self.state.shape[0]
        view = sendbuf[: D1 * ld_send].view(D1, ld_send)
        view[:, :m_loc].copy_(torch.from_numpy(np.ascontiguousarray(self.state[:, anc])))


def _worker(rank, world, port, N, d, seed, conc, u0, mode, out):
    import torch.distributed as dist
    import smcb200
    from oracle import smc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    engine = import_module(smcb200.__name__ + ".engine")
    w, state = _problem(N, d, seed, conc)
    n = N // world
    sl = slice(rank * n, (rank + 1) * n)
    D1 = d + 1
    path = "fused" if mode == "fixed_fused" else "chain"
    mode = "fixed" if mode == "fixed_fused" else mode
    ops = (_OracleFusedOps if path == "fused" else _OracleOps)(smc, w[sl], state[:, sl], N, u0, rank)
    sendbuf = torch.zeros(D1 * N, dtype=torch.float64)
    recvbuf = torch.zeros(D1 * n, dtype=torch.float64)
    state_out = torch.zeros((D1, n), dtype=torch.float64)
    filled = engine.sharded_resample(ops, engine.TorchComm(), N, n, D1, u0, mode, sendbuf, recvbuf, state_out)
    np.savez(os.path.join(out, f"r{rank}.npz"), state=state_out.numpy(), filled=filled)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["fixed", "fixed_fused", "sequential"])
@pytest.mark.parametrize("conc,u0", [(0.3, 0.37), (0.02, 0.0)])
def test_sharded_resample_equals_unsharded(tmp_path, world, mode, conc, u0):
    import torch.multiprocessing as mp
    from oracle import smc
    N, d, seed = 1200, 3, 5
    mp.spawn(_worker, args=(world, _free_port(), N, d, seed, conc, u0, mode, str(tmp_path)), nprocs=world, join=True)
    w, state = _problem(N, d, seed, conc)
    ref_fn = smc.resample_sequential if mode == "sequential" else smc.resample_fixed
    anc, counts, info = ref_fn(w, u0)
    want = state[:, smc.fit_ancestors(anc, N)]
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    got = np.concatenate([p["state"] for p in parts], axis=1)
    assert np.array_equal(got, want)                                          # bit-identical particle set, in order
    assert all(int(p["filled"]) == info["n_filled"] for p in parts)


def test_migration_plan_invariants():
    """Pure integer host logic: every slot is filled exactly once, chunks are contiguous per source."""
    import smcb200
    rs = np.random.RandomState(0)
    for world in (1, 2, 4, 8):
        for trial in range(20):
            n_local = int(rs.randint(1, 50))
            N = n_local * world
            fl = rs.multinomial(N - int(rs.randint(0, min(N, 7) + 1)), rs.dirichlet(np.ones(world)))
            R = N - int(fl.sum())                                             # residual mass in units of 1/N
            q = np.zeros(world, dtype=object)
            cuts = np.sort(rs.randint(0, (R << 62) // N + 1, world - 1)) if world > 1 else np.array([], dtype=object)
            total_q = (R << 62) // N
            edges = [0] + [int(c) for c in cuts] + [total_q]
            q = [edges[i + 1] - edges[i] for i in range(world)]
            plan = smcb200.migration_plan(fl, q, N, n_local, float(rs.uniform()), world)
            send = np.array(plan["send"])
            assert send.sum() == N                                            # clamped / padded to exactly N slots
            assert np.all(send.sum(axis=0) == n_local)                        # every rank receives its n_local slots
            assert [sum(r) for r in plan["send"]] == plan["M"]
            assert plan["O"] == sorted(plan["O"])
