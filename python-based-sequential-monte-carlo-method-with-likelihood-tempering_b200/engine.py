"""Host side of the tempered-SMC sampler loop.

`Engine.run` is the drop-in for the body of the reference driver
(`SMC_example/Micmem_SMC_main.py:98-262`, same loop in
`SMC_methanation/SMC_methanation_main.py:194-418`): prior particles + likelihood + settings in,
posterior particles + tempering schedule + log-evidence out.  The stage-level methods
(`sim_particle`, `temper`, `resample`, `mh_sweep`) keep the reference's script shape usable.

All arithmetic over particles happens in libsmcb200.so (hand-written sm_100a kernels) through
ctypes; torch owns the device buffers and supplies streams and (for sharded runs)
`torch.distributed` collectives.  The host only sees O(1) scalars per stage.

Particle state is one SoA tensor `state[d+1, n_local]` (rows 0..d-1 parameters, row d the
log-likelihood) so resampling moves everything with a single gather.  With `world > 1` particles
are sharded in contiguous blocks; global particle id = rank * n_local + i.
"""
import contextlib
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .settings import Settings

_SCAN = {"sequential": _lib.SCAN_SEQUENTIAL, "fixed": _lib.SCAN_FIXED}
_TWO62 = 1 << 62


# ------------------------------------------------------------------------------------ collectives
class LocalComm:
    """Single-process stand-in (world = 1)."""
    rank, world = 0, 1

    def all_reduce_max(self, t):
        return t

    def all_reduce_sum(self, t):
        return t

    def all_gather_i64(self, t):
        return t.reshape(1, -1)

    def all_to_all(self, out, inp, out_splits, in_splits):
        out.copy_(inp)

    def broadcast(self, t, src):
        return t

    def barrier(self):
        pass


class TorchComm:
    """`torch.distributed` collectives.  NOT the product path: it exists so that the exchange logic of
    `sharded_resample` can be exercised on the CPU (gloo, tests/test_sharding_gloo.py).  An Engine given a TorchComm
    builds an `NcclComm` over the same ranks and uses that."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_i64(self, t):
        out = torch.empty((self.world, t.numel()), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        return out

    def all_to_all(self, out, inp, out_splits, in_splits):
        self.dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits,
                                    group=self.group)

    def exchange_rows(self, send, ld_send, send_counts, recv, ld_recv, recv_counts, rows):
        """CPU twin of smcb_comm_exchange_rows: column ranges of the rows of `send` (viewed as [rows][ld_send]) go to
        the ranks, what arrives lands in column ranges of the rows of `recv` ([rows][ld_recv])."""
        S = send.reshape(-1)[: rows * ld_send].view(rows, ld_send)
        R = recv if recv.dim() == 2 else recv.reshape(-1)[: rows * ld_recv].view(rows, ld_recv)
        so = np.concatenate([[0], np.cumsum(send_counts)]).astype(int)
        ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(int)
        inp = torch.cat([S[:, so[q]:so[q + 1]].reshape(-1) for q in range(self.world)])
        out = torch.empty(rows * int(ro[-1]), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(out, inp, output_split_sizes=[rows * int(c) for c in recv_counts],
                                    input_split_sizes=[rows * int(c) for c in send_counts], group=self.group)
        off = 0
        for q in range(self.world):
            c = int(recv_counts[q])
            R[:, ro[q]:ro[q + 1]].copy_(out[off:off + rows * c].view(rows, c))
            off += rows * c

    def broadcast(self, t, src):
        """src is a rank within this communicator's group (torch.distributed wants the global rank)."""
        g_src = src if self.group is None else self.dist.get_global_rank(self.group, src)
        self.dist.broadcast(t, src=g_src, group=self.group)
        return t

    def barrier(self):
        self.dist.barrier(group=self.group)


class NcclComm:
    """The communicator of the product path: NCCL inside libsmcb200.so (`smcb_comm_init`), one process per GPU.
    Every exchange is enqueued on the engine's CUDA stream by the library itself; torch.distributed is not on the
    data path (it is used once, optionally, to hand rank 0's NCCL unique id to the other ranks).

    The communicator belongs to a library handle, which this object owns and lends to every Engine constructed with
    it (creating an NCCL communicator costs far more than a sampler run, so it is made once per process)."""

    def __init__(self, unique_id, rank, world, device=None):
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device_index = device if isinstance(device, int) else (torch.device(device).index or 0)
        self.rank, self.world = int(rank), int(world)
        self.handle = _lib.p_void()
        rc = self.lib.smcb_create(self.device_index, _lib.C.byref(self.handle))
        if rc != 0:
            raise _lib.SmcbError(rc, self.lib.smcb_last_error(None).decode())
        buf = _lib.C.create_string_buffer(bytes(unique_id), _lib.COMM_ID_BYTES)
        _lib.check(self.handle, self.lib.smcb_comm_init(self.handle, buf, _lib.COMM_ID_BYTES, self.rank, self.world))

    @staticmethod
    def unique_id():
        """128 bytes made by rank 0 and handed to every rank (any transport)."""
        lib = _lib.load()
        buf = _lib.C.create_string_buffer(_lib.COMM_ID_BYTES)
        rc = lib.smcb_comm_unique_id(buf, _lib.COMM_ID_BYTES)
        if rc != 0:
            raise _lib.SmcbError(rc, lib.smcb_last_error(None).decode())
        return buf.raw

    _from_torch = {}

    @classmethod
    def from_torch_distributed(cls, group=None):
        """Builds the communicator over the ranks of an initialised torch.distributed group: rank 0's unique id
        travels through `broadcast_object_list`; nothing else of torch.distributed is used afterwards."""
        import torch.distributed as dist
        key = (id(group), torch.cuda.current_device())
        if key in cls._from_torch:
            return cls._from_torch[key]
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        src = 0 if group is None else dist.get_global_rank(group, 0)
        dist.broadcast_object_list(box, src=src, group=group)
        comm = cls(box[0], rank, world)
        cls._from_torch[key] = comm
        return comm

    # exchanges used by sharded_resample (device tensors in, enqueued on the current stream)
    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def all_gather_i64(self, t):
        t = t.contiguous()
        out = torch.empty((self.world, t.numel()), dtype=t.dtype, device=t.device)
        _lib.check(self.handle, self.lib.smcb_comm_all_gather(self.handle, t.data_ptr(), out.data_ptr(),
                                                             t.numel() * t.element_size(), self._st()))
        return out

    def all_to_all(self, out, inp, out_splits, in_splits):
        sc = np.ascontiguousarray(in_splits, dtype=np.int64)
        rc_ = np.ascontiguousarray(out_splits, dtype=np.int64)
        _lib.check(self.handle, self.lib.smcb_comm_all_to_all_v(self.handle, inp.data_ptr(), sc.ctypes.data,
                                                               out.data_ptr(), rc_.ctypes.data, self._st()))

    def exchange_rows(self, send, ld_send, send_counts, recv, ld_recv, recv_counts, rows):
        sc = np.ascontiguousarray(send_counts, dtype=np.int64)
        rc_ = np.ascontiguousarray(recv_counts, dtype=np.int64)
        _lib.check(self.handle, self.lib.smcb_comm_exchange_rows(self.handle, send.data_ptr(), int(ld_send), sc.ctypes.data,
                                                                recv.data_ptr(), int(ld_recv), rc_.ctypes.data, int(rows),
                                                                self._st()))

    def broadcast(self, t, src):
        _lib.check(self.handle, self.lib.smcb_comm_broadcast(self.handle, t.data_ptr(), t.numel() * t.element_size(),
                                                            int(src), self._st()))
        return t

    def all_reduce_sum(self, t):
        _lib.check(self.handle, self.lib.smcb_comm_all_reduce_f64(self.handle, t.data_ptr(), t.numel(), _lib.OP_SUM,
                                                                 self._st()))
        return t

    def all_reduce_max(self, t):
        _lib.check(self.handle, self.lib.smcb_comm_all_reduce_f64(self.handle, t.data_ptr(), t.numel(), _lib.OP_MAX,
                                                                 self._st()))
        return t

    def barrier(self):
        torch.cuda.current_stream().synchronize()

    def collective_count(self):
        return int(self.lib.smcb_collective_count(self.handle))


def fixed_crossings(s, N, u0q):
    """Thresholds (u0+k)/N, k>=0, at or below the fixed-point prefix s (mirrors resample.cu)."""
    x = s * N
    return 0 if x < u0q else ((x - u0q) >> 62) + 1


def migration_plan(floor_tot, q_tot, N, n_local, u0, world):
    """Where every shard's copies go (pure integer host logic, identical on all ranks).

    floor_tot[r], q_tot[r]: per-shard sums of floor counts / fixed-point residuals.
    Returns dict with, per rank r: carry_q[r], first global slot O[r], number of slots M[r] it fills
    (after clamping to N and padding the tail), and send[r][q] = number of slots rank r sends to q.
    """
    u0q = int(round(u0 * _TWO62)) if not isinstance(u0, int) else u0
    carry_q, O, M = [], [], []
    pre_q = 0
    pre_f = 0
    for r in range(world):
        carry_q.append(pre_q)
        c0 = fixed_crossings(pre_q, N, u0q) if r > 0 else 0
        c1 = fixed_crossings(pre_q + int(q_tot[r]), N, u0q)
        O.append(pre_f + c0)
        M.append(int(floor_tot[r]) + c1 - c0)
        pre_q += int(q_tot[r])
        pre_f += int(floor_tot[r])
    plan = _plan_from_offsets(O, M, N, n_local, world)
    plan.update(carry_q=carry_q, u0q=u0q)
    return plan


# ------------------------------------------------------------------------------------ sharded resampling
def sharded_resample(ops, comm, N, n_local, D1, u0, scan_mode, sendbuf, recvbuf, state_out, timer=None):
    """Cross-shard residual-systematic resampling: every rank ends up with slots
    [rank*n_local, (rank+1)*n_local) of the globally resampled particle set in state_out[D1, n_local].

    The arithmetic over particles is done by `ops` (the CUDA kernels behind the C-ABI in the engine; NumPy twins
    of the oracle in the CPU/gloo tests), the exchanges by `comm`:
        ops.totals() -> int64[2] tensor            (sum of floor counts, sum of fixed-point residuals)
        ops.counts_fixed(carry_q) -> None          counts of this shard given the residual prefix of lower ranks
        ops.counts_sequential(carry[2]) -> (carry_out[2], int64[2] tensor = (floor sum, crossings))
        ops.expand_and_pack(m_loc, send, sendbuf)  ancestors of the first m_loc slots, gathered into one
                                                   contiguous [D1][len] chunk per destination rank
    Returns the number of slots filled before clamping to N."""
    W, rank = comm.world, comm.rank
    timer = timer or (lambda name: contextlib.nullcontext())
    if scan_mode == "fixed" and hasattr(ops, "fused_expand") and hasattr(comm, "exchange_rows"):
        # Device path: totals -> all-gather -> integer plan on the host -> ONE kernel from weights to the rows of the
        # slots this shard fills (smcb_resample_fused with the residual prefix of the lower ranks) -> the column
        # ranges of those rows travel straight into the destination's state (smcb_comm_exchange_rows): no packing by
        # destination, no unpacking.
        with timer("rs_totals"):
            allt = comm.all_gather_i64(ops.totals()).cpu().numpy()
        plan = migration_plan(allt[:, 0], allt[:, 1], N, n_local, u0, W)
        m_loc, send = plan["M"][rank], plan["send"][rank]
        recv = [plan["send"][r][rank] for r in range(W)]
        ld_send = max(m_loc, 1)
        with timer("rs_expand"):
            if m_loc > 0:
                ops.fused_expand(plan["carry_q"][rank], m_loc, sendbuf, ld_send)
        with timer("rs_exchange"):
            comm.exchange_rows(sendbuf, ld_send, send, state_out, state_out.stride(0), recv, D1)
        return plan["filled"]
    if scan_mode == "fixed":
        allt = comm.all_gather_i64(ops.totals()).cpu().numpy()
        plan = migration_plan(allt[:, 0], allt[:, 1], N, n_local, u0, W)
        ops.counts_fixed(plan["carry_q"][rank])
    else:
        # the reference's running sum is inherently serial: shards take turns, passing the carry
        carry = np.array([0.0, u0 * (1.0 / N)], dtype=np.float64)
        carry_t = torch.zeros(2, dtype=torch.float64, device=sendbuf.device)
        tot = None
        for r in range(W):
            if r == rank:
                carry, tot = ops.counts_sequential(carry)
                carry_t.copy_(torch.from_numpy(np.asarray(carry, dtype=np.float64)))
            comm.broadcast(carry_t, src=r)          # r is a rank of the communicator's group
            carry = carry_t.cpu().numpy().copy()
        allt = comm.all_gather_i64(tot).cpu().numpy()   # [W, 2] = (floor sum, crossings)
        O, M, pre = [], [], 0
        for r in range(W):
            O.append(pre)
            M.append(int(allt[r, 0] + allt[r, 1]))
            pre += M[-1]
        plan = _plan_from_offsets(O, M, N, n_local, W)
    m_loc = plan["M"][rank]
    send = plan["send"][rank]
    recv = [plan["send"][r][rank] for r in range(W)]
    if m_loc > 0:
        ops.expand_and_pack(m_loc, send, sendbuf)
    comm.all_to_all(recvbuf[: D1 * sum(recv)], sendbuf[: D1 * sum(send)], [D1 * c for c in recv],
                    [D1 * c for c in send])
    off = 0
    unpack = getattr(ops, "unpack", None)
    for r in range(W):
        if recv[r]:
            if unpack is not None:      # device: a 2-D copy of the library
                unpack(recvbuf, D1 * off, recv[r], state_out, off)
            else:
                state_out[:, off:off + recv[r]].copy_(recvbuf[D1 * off: D1 * (off + recv[r])].view(D1, recv[r]))
            off += recv[r]
    return plan["filled"]


class _DeviceShardOps:
    """`ops` of sharded_resample on the GPU: the C-ABI kernels on the engine's buffers."""

    def __init__(self, eng):
        self.e = eng

    def totals(self):
        e = self.e
        tot = e.icnt[4:6]
        e._ck(e.lib.smcb_resample_totals(e.h, e.w.data_ptr(), e.n, e.N, tot.data_ptr(), e._stream))
        return tot

    def counts_fixed(self, carry_q):
        e = self.e
        e._ck(e.lib.smcb_resample_counts(e.h, e.w.data_ptr(), e.n, e.N, 0.0 if e._u0 is None else e._u0,
                                         _lib.SCAN_FIXED, None, carry_q, e.id_offset, e.counts.data_ptr(),
                                         e.icnt[4:6].data_ptr(), e._stream))

    def counts_sequential(self, carry):
        e = self.e
        carry = np.ascontiguousarray(carry, dtype=np.float64)
        tot = e.icnt[4:6]
        e._ck(e.lib.smcb_resample_counts(e.h, e.w.data_ptr(), e.n, e.N, e._u0, _lib.SCAN_SEQUENTIAL,
                                         carry.ctypes.data, 0, e.id_offset, e.counts.data_ptr(), tot.data_ptr(),
                                         e._stream))
        return carry, tot

    def fused_expand(self, carry_q, m_loc, sendbuf, ld_send):
        """Counts, offsets, ancestors and the gather of this shard's first m_loc output slots in one kernel; the
        rows land in sendbuf viewed as [d+1][ld_send]."""
        e = self.e
        e._ck(e.lib.smcb_resample_fused(e.h, None, e.w.data_ptr(), e.n, e.N, carry_q, e.id_offset, m_loc, None, 0.0,
                                        None, 0.0 if e._u0 is None else e._u0, e.state.data_ptr(), e.n, e.d + 1,
                                        sendbuf.data_ptr(), ld_send, e.anc.data_ptr(), None, e.icnt[6:7].data_ptr(),
                                        e._stream))

    def unpack(self, recvbuf, src_off, width, state_out, dst_off):
        e = self.e
        e._ck(e.lib.smcb_copy_rows(e.h, recvbuf[src_off:].data_ptr(), width, state_out[:, dst_off:].data_ptr(),
                                   state_out.stride(0), width, e.d + 1, e._stream))

    def expand_and_pack(self, m_loc, send, sendbuf):
        e = self.e
        D1 = e.d + 1
        e._ck(e.lib.smcb_ancestors(e.h, e.counts.data_ptr(), e.n, m_loc, e.anc.data_ptr(), e.icnt[6:7].data_ptr(),
                                   e._stream))
        off = 0
        for q, cnt in enumerate(send):   # one contiguous [d+1][len] chunk per destination
            if cnt:
                e._ck(e.lib.smcb_gather(e.h, e.state.data_ptr(), e.n, e.anc[off:].data_ptr(), cnt, D1,
                                        sendbuf[D1 * off:].data_ptr(), cnt, e._stream))
                off += cnt


# ------------------------------------------------------------------------------------ results
@dataclass
class StageRecord:
    step: int
    gamma: float
    ess: float
    max_lk: float
    n_backoff: int
    n_mh: int
    moved: int
    log_evidence: float
    mhstep_ratio: float
    filled: int


_PINNED = {}


def _to_host(t):
    """Device tensor -> fresh NumPy array through a cached pinned staging buffer (a pageable D2H runs at a fifth
    of the pinned rate; the staging buffer is reused, the returned array is the caller's own)."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    buf = _PINNED.get(t.dtype)
    if buf is None or buf.numel() < t.numel():
        buf = torch.empty(max(t.numel(), 1), dtype=t.dtype).pin_memory()
        _PINNED[t.dtype] = buf
    stage = buf[: t.numel()].view(t.shape)
    stage.copy_(t)
    return stage.numpy().copy()


class Result:
    """What a run returns.  Scalars and per-stage records are host values; the posterior particles and their
    log-likelihoods stay on the device (a snapshot of this shard's state) and are copied to the host the first
    time `particles` / `lk` is read."""

    def __init__(self, state_dev, d, betas, ess, log_evidence, n_moved, n_mh, stages, n_eval, n_eval_cut,
                 n_eval_reference, seconds, reached_one, ancestors=None):
        self._state, self._d = state_dev, d
        self._particles = self._lk = None
        self.betas, self.ess, self.log_evidence = betas, ess, log_evidence
        self.n_moved, self.n_mh, self.stages = n_moved, n_mh, stages
        self.n_eval = n_eval                      # particle log-likelihood evaluations requested on the device (global)
        self.n_eval_cut = n_eval_cut              # of those, proposals whose rejection was proven before the last observation
        self.n_eval_reference = n_eval_reference  # what the reference would have evaluated: N * (1 + sweeps)
        self.seconds = seconds                    # device time first sweep -> end of gamma=1 block
        self.reached_one = reached_one
        self.ancestors = ancestors if ancestors is not None else []

    @property
    def particles_device(self):
        """[d, n_local] device tensor (SoA) of this shard's posterior particles."""
        return self._state[: self._d]

    @property
    def particles(self):
        """[n_local, d] posterior particles on the host (this rank's shard when sharded)."""
        if self._particles is None:
            self._particles = _to_host(self._state[: self._d].t())
        return self._particles

    @property
    def lk(self):
        """[n_local] log-likelihoods of the posterior particles on the host."""
        if self._lk is None:
            self._lk = _to_host(self._state[self._d])
        return self._lk

    def __repr__(self):
        return (f"Result(stages={len(self.betas)}, reached_one={self.reached_one}, log_evidence={self.log_evidence!r}, "
                f"n_eval={self.n_eval}, n_eval_cut={self.n_eval_cut}, seconds={self.seconds!r})")


# ------------------------------------------------------------------------------------ handle pool
# cudaMalloc / cudaFree of the library's scratch (a few hundred MB at 2^20 particles) cost more than a whole
# sampler run, so a closed Engine parks its handle here and the next Engine on that device takes it over
# (smcb_reserve only ever grows the scratch).  Handles are destroyed at interpreter exit.
_POOL = {}


def _acquire_handle(lib, device_index):
    free = _POOL.get(device_index)
    if free:
        return free.pop()
    h = _lib.p_void()
    rc = lib.smcb_create(device_index, _lib.C.byref(h))
    if rc != 0:
        raise _lib.SmcbError(rc, lib.smcb_last_error(None).decode())
    return h


def _release_handle(device_index, h):
    _POOL.setdefault(device_index, []).append(h)


def release_pool():
    """Destroy every parked handle (frees the device scratch)."""
    lib = _lib.load()
    for free in _POOL.values():
        while free:
            lib.smcb_destroy(free.pop())


import atexit  # noqa: E402

atexit.register(release_pool)


# ------------------------------------------------------------------------------------ engine
def auto_mm_budget(n_local, n_ex, sm_count):
    """Attempted RK steps after which the bulk MM_PROGRESS kernel hands a solve to the tail kernel, by size
    (`Settings.mm_budget = 0`).  While the solves of a sweep outnumber the lanes of the bulk kernel (6 blocks of 128
    threads per SM) many times over, a long budget keeps solves of a few hundred steps out of the tail kernel (2^20
    particles: 512 -> 72.0 ms per run, 256 -> 73.7, 128 -> 74.5); when every solve has a lane to itself the bulk kernel
    lasts as long as its slowest solve is allowed to and the tail kernel takes the same steps faster (N = 1000: 16 ->
    13.9 ms, 64 -> 14.7, 512 -> 17.3; N = 2^18: 128 -> 44.3, 512 -> 48.0; profiles/budget_by_size_r02.log).  Results do
    not depend on it beyond rounding."""
    waves = n_local * n_ex / float(sm_count * 6 * 128)
    if waves <= 1.0:
        return 32
    if waves <= 20.0:
        return 128
    if waves <= 40.0:
        return 256
    return 512


class Engine:
    def __init__(self, likelihood, prior, settings=None, device=None, comm=None):
        if not torch.cuda.is_available():
            raise RuntimeError("a CUDA device is required: the sampler path has no CPU fallback")
        self.lib = _lib.load()
        self.cfg = (settings or Settings()).validate()
        self.lik, self.prior = likelihood, prior
        self.d = prior.d
        if likelihood.d != self.d:
            raise ValueError(f"prior has {self.d} parameters, likelihood expects {likelihood.d}")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.set_device(self.device)
        if isinstance(comm, TorchComm):
            comm = NcclComm.from_torch_distributed(comm.group) if comm.world > 1 else None
        self.comm = comm or LocalComm()
        if isinstance(self.comm, NcclComm):
            if self.comm.device_index != self.device.index:
                raise ValueError("the communicator was made on another device")
            self.h, self._own_handle = self.comm.handle, False      # the communicator lives in this handle
        else:
            self.h, self._own_handle = _acquire_handle(self.lib, self.device.index), True
        n_local = self.cfg.n_particle // max(self.comm.world, 1)
        self.mm_budget = self.cfg.mm_budget or self._auto_mm_budget(n_local, likelihood)
        for key, val in ((_lib.PARAM_MM_BUDGET, self.mm_budget), (_lib.PARAM_MM_REFILL_MIN, self.cfg.mm_refill_min),
                         (_lib.PARAM_MM_PATIENCE, self.cfg.mm_patience), (_lib.PARAM_MM_CHUNK, self.cfg.mm_chunk),
                         (_lib.PARAM_MM_TAIL_WARPS, self.cfg.mm_tail_warps)):
            self._ck(self.lib.smcb_set_param(self.h, key, float(val)))
        likelihood.upload(self.lib, self.h)
        N, W = self.cfg.n_particle, self.comm.world
        if N % W != 0:
            raise ValueError("n_particle must be divisible by the number of ranks")
        self.N, self.n = N, N // W
        self.id_offset = self.comm.rank * self.n
        # a shard may have to expand up to N copies when it holds all the weight
        self.cap = self.n if W == 1 else N
        self._ck(self.lib.smcb_reserve(self.h, max(self.cap, self.n), self.d))
        dev, f64 = self.device, torch.float64
        D1 = self.d + 1
        self.state = torch.zeros((D1, self.n), dtype=f64, device=dev)
        self.state2 = torch.zeros((D1, self.n), dtype=f64, device=dev)
        self.prop = torch.zeros((self.d, self.n), dtype=f64, device=dev)
        self.lk2 = torch.zeros(self.n, dtype=f64, device=dev)
        self.w = torch.zeros(self.n, dtype=f64, device=dev)
        self.lkmin = torch.zeros(self.n, dtype=f64, device=dev)
        self.inbox = torch.zeros(self.n, dtype=torch.uint8, device=dev)
        self.moved = torch.zeros(self.n, dtype=torch.uint8, device=dev)
        self.counts = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.anc = torch.zeros(self.cap, dtype=torch.int32, device=dev)
        self.scal = torch.zeros(128, dtype=f64, device=dev)      # [0]=max, [1]=accepted sum_w, [2:]=tempering sums
        self.filled_hist = torch.zeros(max(self.cfg.itr_max, 1) + 1, dtype=torch.int64, device=dev)
        # sweep block (smcb_moments_merged): [0:4] MH counters over all ranks, [4:4+d] mean, then M2 [d*d], then the
        # device-side proposal factor F [d*d]; its first part comes back to the host once per sweep
        dd = self.d * self.d
        self.blk = torch.zeros(4 + self.d + 2 * dd, dtype=f64, device=dev)
        self._h_blk = torch.zeros(4 + self.d + dd, dtype=f64).pin_memory()
        self._h_scal = torch.zeros(128, dtype=f64).pin_memory()
        self._w_cov_c = np.ascontiguousarray(self.cfg.w_cov(self.d), dtype=np.float64)
        self.icnt = torch.zeros(8, dtype=torch.int64, device=dev)  # [0:4] MH counters, [4:6] totals, [6] filled
        self.sendbuf = None
        self._w_cov = None
        self.prof = None          # name -> list of (start, end) CUDA events when profiling is on
        self._low = np.ascontiguousarray(prior.low)
        self._high = np.ascontiguousarray(prior.high)
        self.dlp = None           # log prior ratio of the current proposals (priors with normal components only)
        if getattr(prior, "has_normal", False):
            if self.cfg.fused_sweeps > 0:
                raise NotImplementedError("several sweeps per call support uniform priors only")
            self.dlp = torch.zeros(self.n, dtype=f64, device=dev)

    # -------------------------------------------------------------------------------- helpers
    def _ck(self, rc):
        _lib.check(self.h, rc)

    @property
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        """Park the library handle (with its device scratch) for the next Engine on this device."""
        if getattr(self, "h", None):
            if self.lik.model_id == _lib.MODEL_USER:      # the callback must not outlive the likelihood object
                self.lib.smcb_set_user_likelihood(self.h, None, None)
            if self._own_handle:
                _release_handle(self.device.index, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return int(self.lib.smcb_launch_count(self.h))

    def loglik_stats(self):
        """int64[16] work counters of the MM progress-curve kernels (see smcb_loglik_stats)."""
        out = np.zeros(_lib.N_STATS, dtype=np.int64)
        self._ck(self.lib.smcb_loglik_stats(self.h, out.ctypes.data))
        return out

    def kernel_profile(self, on=None):
        """Per-kernel device time of the MM_PROGRESS sweeps (CUDA events inside the library).
        on=True/False switches the recording; with on=None returns (bulk ms, tail ms, sweeps) since the
        last read."""
        if on is not None:
            self._ck(self.lib.smcb_set_param(self.h, _lib.PARAM_PROFILE, 1.0 if on else 0.0))
            return None
        out = np.zeros(3, dtype=np.float64)
        self._ck(self.lib.smcb_profile_read(self.h, out.ctypes.data))
        return float(out[0]), float(out[1]), int(out[2])

    # -------------------------------------------------------------------------------- device timers
    def enable_profiling(self, on=True):
        """Bracket every kernel group with CUDA events on the launching stream (bench.py's roofline)."""
        self.prof = {} if on else None

    @contextlib.contextmanager
    def _timed(self, name):
        if self.prof is None:
            yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        self.prof.setdefault(name, []).append((e0, e1))

    def profile_summary(self):
        """{name: (launch groups, total ms)}; synchronises."""
        torch.cuda.synchronize(self.device)
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (self.prof or {}).items()}

    def profile_events(self, name):
        """Per-launch-group device times [ms] of one kernel group, in launch order; synchronises."""
        torch.cuda.synchronize(self.device)
        return [a.elapsed_time(b) for a, b in (self.prof or {}).get(name, [])]

    @property
    def theta(self):
        return self.state[: self.d]

    @property
    def lk(self):
        return self.state[self.d]

    # -------------------------------------------------------------------------------- particles
    def set_particles(self, particles):
        """particles: [n_local, d] (reference layout, AoS) host or device array of this shard."""
        p = torch.as_tensor(particles, dtype=torch.float64)
        if p.shape != (self.n, self.d):
            raise ValueError(f"expected particles of shape {(self.n, self.d)}, got {tuple(p.shape)}")
        self.state[: self.d].copy_(p.to(self.device, non_blocking=True).t())

    def sample_prior(self, seed=None):
        """Prior sample drawn on the device with Philox (keyed by global particle id): uniform components by
        smcb_sample_uniform_box, normal ones as mu + sigma*z with the z of smcb_philox_draws (stage 0xFFFFFFFE)."""
        seed = self.cfg.seed if seed is None else seed
        low, high = self._low, self._high
        if self.dlp is not None:        # finite stand-ins for the unbounded components; overwritten below
            low = np.where(np.isfinite(low), low, 0.0)
            high = np.where(np.isfinite(high), high, 1.0)
        self._ck(self.lib.smcb_sample_uniform_box(self.h, self.state.data_ptr(), self.n, self.n, self.d,
                                                  low.ctypes.data, high.ctypes.data, seed,
                                                  self.id_offset, self._stream))
        if self.dlp is not None:
            z = torch.empty((self.n, self.d), dtype=torch.float64, device=self.device)
            self._ck(self.lib.smcb_philox_draws(self.h, self.n, self.d, seed, self.id_offset, 0xFFFFFFFE, 0,
                                                z.data_ptr(), None, self._stream))
            for k, (kind, a, b) in enumerate(self.prior.dists):
                if kind == "normal":
                    self.state[k].copy_(a + b * z[:, k])

    def particles(self):
        """[n_local, d] tensor (AoS copy of the current shard)."""
        return self.state[: self.d].t().contiguous()

    # -------------------------------------------------------------------------------- checkpoint
    def state_dict(self):
        """Everything needed to continue the last `run(max_stages=...)` of this shard: particles, their
        log-likelihoods, gamma, stage counter, log-evidence, counters and the host RNG state (the device RNG is
        Philox keyed by (seed, particle id, stage, sweep): it has no state to save)."""
        if getattr(self, "_ckpt", None) is None:
            raise RuntimeError("state_dict() is available after run()")
        out = dict(self._ckpt)
        out["particles"] = self.particles().cpu().numpy()
        out["lk"] = self.lk.cpu().numpy().copy()
        out["n_particle"], out["rank"], out["world"], out["seed"] = self.N, self.comm.rank, self.comm.world, self.cfg.seed
        return out

    def load_state_dict(self, sd):
        if sd["n_particle"] != self.N or sd["world"] != self.comm.world or sd["rank"] != self.comm.rank:
            raise ValueError("checkpoint was taken with a different particle count or sharding")
        self.set_particles(sd["particles"])
        self.lk.copy_(torch.as_tensor(sd["lk"], dtype=torch.float64))

    # -------------------------------------------------------------------------------- K1
    def loglik_into(self, theta, lk_out, active=None, lkmin=None):
        """lk_out[i] = log-likelihood of theta[:, i] (active: byte mask of the particles to evaluate;
        lkmin: early-rejection thresholds, see smcb_loglik_bounded)."""
        n = theta.shape[1]
        with self._timed("loglik"):
            self._ck(self.lib.smcb_loglik_bounded(self.h, self.lik.model_id, theta.data_ptr(), theta.stride(0), n,
                                                  self.d, active.data_ptr() if active is not None else None,
                                                  lkmin.data_ptr() if lkmin is not None else None,
                                                  lk_out.data_ptr(), self._stream))

    def sim_particle(self, particle=None):
        """Reference surface (`sim_particle(particle) -> llk`, Micmem_likelihood.py:79-92): evaluates
        all local particles; returns the log-likelihood tensor (device)."""
        if particle is not None:
            self.set_particles(particle)
        self.loglik_into(self.theta, self.lk)
        return self.lk

    # -------------------------------------------------------------------------------- K2
    def _auto_mm_budget(self, n_local, likelihood):
        """Deferral budget of the progress-curve likelihood by size (auto_mm_budget)."""
        t = getattr(likelihood, "t", None)
        n_ex = int(t.shape[0]) if isinstance(t, np.ndarray) and t.ndim == 2 else 1
        return auto_mm_budget(n_local, n_ex, torch.cuda.get_device_properties(self.device).multi_processor_count)

    def temper(self, gamma_old):
        """Next gamma by the configured rule.  Returns dict(gamma_new, gm, ess, sum_w, max_lk, n_backoff)
        and leaves max in scal[0], the accepted sum_w in scal[1]."""
        cfg, N = self.cfg, self.N
        st = self._stream
        sums = self.scal[2:2 + 6 * _lib.MAX_CAND]

        def eval_batch(gms):
            """Global max and sums for up to 3*MAX_CAND increments: per-shard reductions, ONE exchange (all-gather of
            the shard rows, logsumexp merge on the device: smcb_temper_eval), one D2H."""
            g = np.ascontiguousarray(gms, dtype=np.float64)
            with self._timed("temper"):
                self._ck(self.lib.smcb_temper_eval(self.h, self.lk.data_ptr(), self.n, g.ctypes.data, len(g),
                                                   self.scal.data_ptr(), st))
            k = 2 + 2 * len(g)
            self._h_scal[:k].copy_(self.scal[:k], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            host = self._h_scal[:k].numpy()
            return float(host[0]), host[2:].copy()

        if cfg.temper_rule == "backoff":
            # candidate list of the reference's geometric back-off (Micmem_SMC_main.py:111-141)
            g_new = gamma_old + cfg.d_gamma_max
            if g_new > 1.0:
                g_new = 1.0
            cands = []
            for _ in range(cfg.gm_reduction_itr):
                cands.append(g_new)
                g_new = (g_new - gamma_old) * cfg.gm_reduction_rate + gamma_old
            g_after_last = g_new
            k_acc, max_lk = None, None
            per_round = 3 * cfg.cand_batch     # the reference's run needs 0..17 back-offs: one rendezvous per stage
            for b0 in range(0, len(cands), per_round):
                batch = cands[b0:b0 + per_round]
                max_lk, s = eval_batch([g - gamma_old for g in batch])
                for k in range(len(batch)):
                    s1, s2 = float(s[2 * k]), float(s[2 * k + 1])
                    ess = 1.0 / (s2 / (s1 * s1)) / N
                    if ess > cfg.ess_limit:
                        k_acc = b0 + k
                        break
                if k_acc is not None:
                    break
            if k_acc is None:
                # the reference keeps the weights of the last tested increment but a gamma reduced once more
                k, k_acc = len(batch) - 1, len(cands) - 1
                gamma_new, n_back = g_after_last, len(cands)
            else:
                k = k_acc - b0
                gamma_new, n_back = cands[k_acc], k_acc
            gm = cands[k_acc] - gamma_old
            s1, s2 = float(s[2 * k]), float(s[2 * k + 1])
            ess = 1.0 / (s2 / (s1 * s1)) / N
            self.scal[1:2].copy_(self.scal[2 + 2 * k: 3 + 2 * k])
            return dict(gamma_new=gamma_new, gm=gm, ess=ess, sum_w=s1, max_lk=max_lk, n_backoff=n_back)

        # bisection on ESS/N = ess_limit (north_star's alternative rule)
        hi = 1.0 - gamma_old
        max_lk, s = eval_batch([hi])

        def ess_of(s, k=0):
            return (s[2 * k] * s[2 * k]) / s[2 * k + 1] / N

        if ess_of(s) >= cfg.ess_limit:
            gm, gamma_new = hi, 1.0
        else:
            lo = 0.0
            for _ in range(cfg.bisect_iters):
                mid = 0.5 * (lo + hi)
                _, s = eval_batch([mid])
                if ess_of(s) >= cfg.ess_limit:
                    lo = mid
                else:
                    hi = mid
                if hi - lo <= 1e-12:
                    break
            gm = lo
            _, s = eval_batch([gm])
            gamma_new = gamma_old + gm
        self.scal[1:2].copy_(sums[0:1])
        return dict(gamma_new=gamma_new, gm=gm, ess=float(ess_of(s)), sum_w=float(s[0]), max_lk=max_lk, n_backoff=0)

    # -------------------------------------------------------------------------------- K3
    def resample(self, gm, u0, weights=None):
        """Residual-systematic resampling of the current state with increment `gm` (weights
        exp((lk-max)*gm)/sum_w from scal[0], scal[1]) or explicit normalised `weights`.
        Returns the number of slots filled before clamping; new state replaces the old."""
        st, lib, h = self._stream, self.lib, self.h
        mode = _SCAN[self.cfg.scan_mode]
        self._u0 = float(u0)
        tot = self.icnt[4:6]
        filled_t = self.icnt[6:7]
        D1, W = self.d + 1, self.comm.world
        if W == 1 and mode == _lib.SCAN_FIXED:
            # the whole of K3 in one kernel: weights, counts, offsets, ancestors and the gather of the state
            w_ptr = None
            if weights is not None:
                self.w.copy_(torch.as_tensor(weights, dtype=torch.float64))
                w_ptr = self.w.data_ptr()
            with self._timed("resample_fused"):
                self._ck(lib.smcb_resample_fused(h, self.lk.data_ptr(), w_ptr, self.n, self.n, 0, 0, self.n,
                                                 self.scal.data_ptr(), gm, self.scal[1:].data_ptr(), u0,
                                                 self.state.data_ptr(), self.n, D1, self.state2.data_ptr(), self.n,
                                                 self.anc.data_ptr(), None, filled_t.data_ptr(), st))
            self.state, self.state2 = self.state2, self.state
            return None   # filled count stays on the device (icnt[6]); read lazily
        if weights is not None:
            self.w.copy_(torch.as_tensor(weights, dtype=torch.float64))
        else:
            with self._timed("weights"):
                self._ck(lib.smcb_weights(h, self.lk.data_ptr(), self.n, self.scal.data_ptr(), gm,
                                          self.scal[1:].data_ptr(), self.w.data_ptr(), st))
        if W == 1:
            with self._timed("resample_scan"):
                self._ck(lib.smcb_resample_counts(h, self.w.data_ptr(), self.n, self.N, u0, mode, None, 0, 0,
                                                  self.counts.data_ptr(), tot.data_ptr(), st))
                self._ck(lib.smcb_ancestors(h, self.counts.data_ptr(), self.n, self.n, self.anc.data_ptr(),
                                            filled_t.data_ptr(), st))
            with self._timed("resample_gather"):
                self._ck(lib.smcb_gather(h, self.state.data_ptr(), self.n, self.anc.data_ptr(), self.n, D1,
                                         self.state2.data_ptr(), self.n, st))
            self.state, self.state2 = self.state2, self.state
            return None   # filled count stays on the device (icnt[6]); read lazily
        # ---- sharded: cross-GPU exclusive scan of shard totals, then all-to-all migration ----
        if self.sendbuf is None:
            self.sendbuf = torch.empty(D1 * self.cap, dtype=torch.float64, device=self.device)
            self.recvbuf = torch.empty(D1 * self.n, dtype=torch.float64, device=self.device)
        with self._timed("resample_sharded"):
            filled = sharded_resample(_DeviceShardOps(self), self.comm, self.N, self.n, D1, u0,
                                      "fixed" if mode == _lib.SCAN_FIXED else "sequential",
                                      self.sendbuf, self.recvbuf, self.state2, timer=self._timed)
        self.state, self.state2 = self.state2, self.state
        return filled

    # -------------------------------------------------------------------------------- K4
    def _launch_moments(self, with_counts=False):
        """Enqueue the merged moment reduction of the current particles (smcb_moments_merged): shard-local two-pass
        moments, ONE all-gather that also carries the MH counters of the sweep just finished, merge and device-side
        proposal factor.  Leaves counters | mean | M2 | F in `self.blk`."""
        with self._timed("moments"):
            self._ck(self.lib.smcb_moments_merged(self.h, self.state.data_ptr(), self.n, self.n, self.d, self.N,
                                                  self.icnt.data_ptr() if with_counts else None,
                                                  self._w_cov_c.ctypes.data, self.blk.data_ptr(), self._stream))

    def _read_block(self, with_moments):
        """Counters (and, for the host-side factor, mean and M2) of the last merged reduction: one D2H, one sync."""
        k = 4 + (self.d + self.d * self.d if with_moments else 0)
        self._h_blk[:k].copy_(self.blk[:k], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._h_blk.numpy()

    def _factor_from_moments(self, cov_sum):
        """cov = np.cov(p_filt.T, bias=True) * w_cov, factorised the way NumPy's legacy multivariate_normal
        does (SVD): x = z @ (sqrt(s)[:,None] * Vt)."""
        d = self.d
        cov = np.array(cov_sum, dtype=np.float64).reshape(d, d) / float(self.N)
        cov = cov * self._w_cov_c
        (u, sv, v) = np.linalg.svd(cov)
        return np.ascontiguousarray(np.sqrt(sv)[:, None] * v), cov

    def proposal_factor(self):
        """Moments of the current particles -> (F, cov) (`Micmem_SMC_main.py:212-215` + the factor of :220),
        NumPy's SVD factor computed on the host."""
        self._launch_moments()
        d = self.d
        blk = self._read_block(True)
        return self._factor_from_moments(blk[4 + d:4 + d + d * d])

    def device_factor(self):
        """The factor the device built in the last merged reduction (Jacobi eigen-decomposition), as a host array."""
        d = self.d
        return self.blk[4 + d + d * d:4 + d + 2 * d * d].cpu().numpy().reshape(d, d)

    def mh_sweep(self, gamma, F, ratio, stage, sweep, Z=None, U=None):
        """One sweep: propose, evaluate in-box proposals, accept.  MH counters accumulate in icnt[0:4].
        F: host factor [d, d], or None to use the one the last merged moment reduction left on the device."""
        st, lib, h, d = self._stream, self.lib, self.h, self.d
        seed = self.cfg.seed
        z_ptr = u_ptr = None
        if Z is not None:
            Zt = torch.as_tensor(Z, dtype=torch.float64).to(self.device).contiguous()
            z_ptr = Zt.data_ptr()
        if U is not None:
            Ut = torch.as_tensor(U, dtype=torch.float64).to(self.device).contiguous()
            u_ptr = Ut.data_ptr()
        with self._timed("propose"):
            if F is None:      # the factor smcb_moments_merged left on the device
                F_dev = self.blk[4 + d + d * d:]
                self._ck(lib.smcb_mh_propose_dev(h, self.state.data_ptr(), self.n, self.n, d, F_dev.data_ptr(), ratio,
                                                 self._low.ctypes.data, self._high.ctypes.data, z_ptr, seed,
                                                 self.id_offset, stage, sweep, self.prop.data_ptr(), self.n,
                                                 self.inbox.data_ptr(), st))
            else:
                F = np.ascontiguousarray(F, dtype=np.float64)
                self._ck(lib.smcb_mh_propose(h, self.state.data_ptr(), self.n, self.n, d, F.ctypes.data, ratio,
                                             self._low.ctypes.data, self._high.ctypes.data, z_ptr, seed,
                                             self.id_offset, stage, sweep, self.prop.data_ptr(), self.n,
                                             self.inbox.data_ptr(), st))
        dlp_ptr = None
        if self.dlp is not None:
            # log p(theta') - log p(theta) over the normal components: pp = exp(px*gamma) * p0_2/p0_1 (main:369)
            self._ck(lib.smcb_prior_logratio(h, self.state.data_ptr(), self.n, self.prop.data_ptr(), self.n, self.n, d,
                                             self.prior.mu.ctypes.data, self.prior.inv2var.ctypes.data,
                                             self.inbox.data_ptr(), self.dlp.data_ptr(), st))
            dlp_ptr = self.dlp.data_ptr()
        lkmin = None
        if self.cfg.early_reject:
            # value below which the accept test below is certain to fail: the likelihood kernel may stop there
            lkmin = self.lkmin
            with self._timed("propose"):
                self._ck(lib.smcb_mh_threshold(h, self.lk.data_ptr(), self.inbox.data_ptr(), self.n, gamma, u_ptr,
                                               dlp_ptr, seed, self.id_offset, stage, sweep, lkmin.data_ptr(), st))
        self.loglik_into(self.prop, self.lk2, active=self.inbox, lkmin=lkmin)
        with self._timed("accept"):
            self._ck(lib.smcb_mh_accept(h, self.state.data_ptr(), self.n, self.lk.data_ptr(), self.prop.data_ptr(),
                                        self.n, self.lk2.data_ptr(), self.inbox.data_ptr(), self.n, d, gamma, u_ptr,
                                        dlp_ptr, seed, self.id_offset, stage, sweep, self.moved.data_ptr(),
                                        self.icnt.data_ptr(), st))

    def mh_fused(self, gamma, F, ratio, stage, sweep0, n_sweeps):
        F = np.ascontiguousarray(F, dtype=np.float64)
        with self._timed("mh_fused"):
            self._ck(self.lib.smcb_mh_fused(self.h, self.lik.model_id, self.state.data_ptr(), self.n,
                                            self.lk.data_ptr(), self.n, self.d, F.ctypes.data, ratio,
                                            self._low.ctypes.data, self._high.ctypes.data, gamma, n_sweeps,
                                            self.cfg.seed, self.id_offset, stage, sweep0, self.moved.data_ptr(),
                                            self.icnt.data_ptr(), self._stream))

    # -------------------------------------------------------------------------------- the loop
    def run(self, particles=None, stream=None, keep_ancestors=False, hook=None, max_stages=None, resume=None):
        """Tempered SMC from gamma=0 to gamma=1.

        particles:  [n_local, d] initial (prior) particles, or None to use what `set_particles` /
                    `sample_prior` left on the device.
        stream:     optional object with u0(), normals(N, d), uniforms(N) supplying the random inputs
                    (parity mode); default is one seeded host draw for u0 and Philox on the device.
        max_stages: stop after this many stages (the state can then be saved with `state_dict`).
        resume:     a `state_dict()` of this shard taken at a stage boundary: the run continues from there and
                    ends exactly where the uninterrupted run would have.
        """
        cfg, N, d = self.cfg, self.N, self.d
        if cfg.fused_sweeps > 0 and stream is not None:
            raise ValueError("several sweeps per call (fused_sweeps > 0) draw their random inputs on the device; "
                             "supplied random inputs need fused_sweeps = 0")
        if particles is not None:
            self.set_particles(particles)
        host_rng = np.random.RandomState(cfg.seed)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        gamma_old, logZ, first_step = 0.0, 0.0, 1
        stages, ancestors = [], []
        if resume is None:
            self.sim_particle()
            n_eval, n_sweeps_total = N, 0    # evaluations requested (global): N + in-box proposals of every sweep
            n_cut = 0                        # of those, proposals rejected early (reported -inf before all observations)
        else:
            self.load_state_dict(resume)
            host_rng.set_state(resume["host_rng_state"])
            gamma_old, logZ, first_step = float(resume["gamma"]), float(resume["log_evidence"]), int(resume["step"]) + 1
            n_eval, n_cut, n_sweeps_total = int(resume["n_eval"]), int(resume["n_eval_cut"]), int(resume["n_sweeps"])
            stages = list(resume["stages"])
        reached = resume is not None and gamma_old >= 1.0      # checkpoint taken after the last stage: nothing to do
        first_step = cfg.itr_max if reached else first_step
        for step in range(first_step, cfg.itr_max):
            if max_stages is not None and step - first_step >= max_stages:
                break
            t = self.temper(gamma_old)
            gamma_new, gm = t["gamma_new"], t["gm"]
            logZ += math.log(t["sum_w"] / N) + gm * t["max_lk"]
            u0 = float(stream.u0()) if stream is not None else float(host_rng.rand())
            filled = self.resample(gm, u0)
            if keep_ancestors and self.comm.world == 1:
                ancestors.append(self.anc[: self.n].cpu().numpy().astype(np.int64))
            self._ck(self.lib.smcb_zero(self.h, self.moved.data_ptr(), self.n, self._stream))
            self._ck(self.lib.smcb_zero(self.h, self.icnt.data_ptr(), 32, self._stream))
            ratio = 1.0
            if gamma_new >= 1.0:
                n_mh, r_th = cfg.ad_mhstep_num, cfg.r_threshold_f
            else:
                n_mh, r_th = cfg.mhstep_num, cfg.r_threshold
            n_run, moved, stage_evals, stage_cut = 0, 0, 0, 0
            # Host-side factor (NumPy's SVD, reproduces the reference's proposals for given normals) in parity mode
            # and for the fused kernel, which takes a host factor; the device's own Jacobi factor otherwise.
            host_factor = cfg.factor == "host" or (cfg.factor == "auto" and (stream is not None or hook is not None))
            if cfg.fused_sweeps > 0 and self.lik.model_id != _lib.MODEL_KINETIC_RK:
                # any model: batches of sweeps with the covariance refreshed every sweep (smcb_mh_sweeps); the
                # reference's early-exit and step-halving rules act between batches
                self._launch_moments()
                done = 0
                while done < n_mh:
                    k = min(cfg.fused_sweeps, n_mh - done)
                    with self._timed("mh_sweeps"):
                        self._ck(self.lib.smcb_mh_sweeps(
                            self.h, self.lik.model_id, self.state.data_ptr(), self.n, self.lk.data_ptr(), self.n, d,
                            self.N, self._w_cov_c.ctypes.data, ratio, self._low.ctypes.data, self._high.ctypes.data,
                            gamma_new, k, 1 if cfg.early_reject else 0, cfg.seed, self.id_offset, step, done,
                            self.prop.data_ptr(), self.n, self.lk2.data_ptr(), self.lkmin.data_ptr(),
                            self.inbox.data_ptr(), self.moved.data_ptr(), self.icnt.data_ptr(), self.blk.data_ptr(),
                            self._stream))
                    done += k
                    n_run += k
                    blk = self._read_block(False)
                    moved, stage_evals, stage_cut = int(blk[1]), int(blk[2]), int(blk[3])
                    if cfg.early_exit and moved > r_th * N:
                        break
                    if moved < cfg.r_threshold_min * N:
                        ratio = ratio * 0.5
            elif cfg.fused_sweeps > 0:
                F, _ = self.proposal_factor()
                done = 0
                while done < n_mh:
                    k = min(cfg.fused_sweeps, n_mh - done)
                    self.mh_fused(gamma_new, F, ratio, step, done, k)
                    done += k
                    n_run += k
                    # counters of the batch and the moments the next batch needs: one exchange, one rendezvous
                    self._launch_moments(with_counts=True)
                    blk = self._read_block(True)
                    moved, stage_evals, stage_cut = int(blk[1]), int(blk[2]), int(blk[3])
                    if cfg.early_exit and moved > r_th * N:
                        break
                    # the reference's step-size rule (Micmem_SMC_main.py:247-248), applied once per fused batch:
                    # with fused_sweeps = 1 this is the reference's per-sweep rule exactly
                    if moved < cfg.r_threshold_min * N:
                        ratio = ratio * 0.5
                    if done < n_mh:
                        F, _ = self._factor_from_moments(blk[4 + d:4 + d + d * d])
            else:
                # One exchange and one host<->device rendezvous per sweep: after a sweep's accept kernel the merged
                # reduction gathers its counters TOGETHER with the moments the next sweep needs (if the early-exit
                # rule then stops the stage those moments were wasted: ~20 us).
                dd = d + d * d
                self._launch_moments()
                if host_factor:
                    blk = self._read_block(True)
                for j in range(n_mh):
                    F = cov = None
                    if host_factor:
                        F, cov = self._factor_from_moments(blk[4 + d:4 + dd])
                    Z = stream.normals(N, d) if stream is not None else None
                    U = stream.uniforms(N) if stream is not None else None
                    if Z is not None and self.comm.world > 1:
                        lo = self.id_offset
                        Z, U = Z[lo:lo + self.n], U[lo:lo + self.n]
                    if hook is not None:
                        hook("sweep", step=step, j=j, engine=self, F=F, cov=cov, gamma=gamma_new, ratio=ratio)
                    self.mh_sweep(gamma_new, F, ratio, step, j, Z, U)
                    n_run += 1
                    self._launch_moments(with_counts=True)
                    blk = self._read_block(host_factor)
                    moved, stage_evals, stage_cut = int(blk[1]), int(blk[2]), int(blk[3])
                    if cfg.early_exit and moved > r_th * N:
                        break
                    if moved < cfg.r_threshold_min * N:
                        ratio = ratio * 0.5
            n_eval += stage_evals
            n_cut += stage_cut
            n_sweeps_total += n_run
            if filled is None:
                self.filled_hist[step].copy_(self.icnt[6])   # read once, after the run
                filled = -1
            stages.append(StageRecord(step, gamma_new, t["ess"], t["max_lk"], t["n_backoff"], n_run, moved, logZ,
                                      ratio, filled))
            if hook is not None:
                hook("stage", step=step, engine=self, gamma=gamma_new)
            if gamma_new == 1.0:
                reached = True
                break
            gamma_old = gamma_new
        ev1.record()
        ev1.synchronize()
        if any(sr.filled < 0 for sr in stages):
            fh = self.filled_hist.cpu().numpy()
            for sr in stages:
                if sr.filled < 0:
                    sr.filled = int(fh[sr.step])
        secs = ev0.elapsed_time(ev1) * 1e-3
        self._ckpt = dict(gamma=gamma_old if not reached else 1.0, log_evidence=logZ,
                          step=stages[-1].step if stages else 0, n_eval=n_eval, n_eval_cut=n_cut,
                          n_sweeps=n_sweeps_total, stages=list(stages), host_rng_state=host_rng.get_state())
        return Result(state_dev=self.state.clone(), d=d,
                      betas=[s.gamma for s in stages], ess=[s.ess for s in stages], log_evidence=logZ,
                      n_moved=[s.moved for s in stages], n_mh=[s.n_mh for s in stages], stages=stages,
                      n_eval=n_eval, n_eval_cut=n_cut, n_eval_reference=N * (1 + n_sweeps_total),
                      seconds=secs, reached_one=reached, ancestors=ancestors)


def _plan_from_offsets(O, M, N, n_local, world):
    filled = sum(M)
    O, M = list(O), list(M)
    for r in range(world):
        lo, hi = min(O[r], N), min(O[r] + M[r], N)
        O[r], M[r] = lo, hi - lo
    if filled < N:
        last = max((r for r in range(world) if M[r] > 0), default=world - 1)
        M[last] += N - (O[last] + M[last])
    send = [[0] * world for _ in range(world)]
    for r in range(world):
        for q in range(world):
            lo = max(O[r], q * n_local)
            hi = min(O[r] + M[r], (q + 1) * n_local)
            send[r][q] = max(0, hi - lo)
    return dict(O=O, M=M, send=send, filled=filled)
