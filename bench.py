#!/usr/bin/env python
"""Benchmark of the tempered-SMC hot path (BASELINE.json metric: particle-loglik evals/s, time-to-beta=1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one complete pass of the hot path over one batch of particles: a whole tempered-SMC
run (first likelihood sweep -> tempering -> residual-systematic resampling -> MH sweeps, repeated
until beta = 1) started from prior particles that are already resident in HBM.  `value` is the
number of per-particle log-likelihood evaluations the device performed divided by the device time
(CUDA events, max over ranks); `time_to_beta1_s` is the average step time.  `e2e` is the same
metric through the public call `smcb200.run(likelihood, prior, host_particles, settings)`: engine
construction, H2D of the particles from pinned host memory, the run, D2H of posterior particles.

Workloads (BASELINE.json configs):
  mm_progress   config 2: Michaelis-Menten progress curves (the reference's six CSVs, 240
                observations), 2^20 particles per GPU, FP64, scipy-RK45-twin arithmetic.  DEFAULT.
  mm_progress_exact  the same problem with the closed-form (Wright omega) progress curves: converged mode, labelled
  mm_rate       config 4 shape: 10 000 rate-law observations, 2^22 particles per GPU (--particles), direct FP32 sum
  mm_rate_suff  config 4, sufficient-statistic form (A(Km), B(Km) tabulated once): 2^23 particles per GPU = 2^26 on 8
  kinetic       config 3: methanation-style reactor, d=5, 30 conditions, RK4 x 50, 2^18 particles
  kinetic_dae   SURVEY 8(f) N3: the reference's transient reactor DAE, 30 conditions, the reference's N = 1000
  kinetic32     config 5: 32-parameter kinetic family, 10 fused MH sweeps per stage, 2^21 particles per GPU
                (2^24 over 8 GPUs; --total-particles 16777216 for the strong-scaling series)
With N > 1 (torchrun) particles are sharded, per-GPU count fixed => "scaling": "weak".

`--impl reference` times the reference's CPU implementation of the same path (the oracle's
restatement: scipy RK45 likelihood fanned out over all host cores + the NumPy sampler-loop body) on
a bounded sample; only rank 0 works.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz")
METRIC = "particle_loglik_evals_per_s"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mm_progress", choices=["mm_progress", "mm_progress_exact", "mm_rate", "mm_rate_suff", "kinetic", "kinetic32", "kinetic_dae"])
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (0 = workload default)")
    ap.add_argument("--total-particles", type=int, default=0,
                    help="strong scaling: this many particles in total, split evenly over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------------------------ workloads
def make_workload(pkg, name, n_per_gpu):
    if name.startswith("kinetic"):
        # synthetic operating conditions / observations: fixture written by tests/golden/make_kinetic_fixture.py
        kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
    if name == "mm_progress":
        g = np.load(GOLDEN)
        lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
        prior = pkg.UniformBox([0, 0, 0], [10, 10, 10], names=["Vmax", "Km", "sigma"])
        n = n_per_gpu or (1 << 20)
        cfg = dict()
        desc = "Michaelis-Menten tempered SMC, 6x40 progress-curve observations (reference CSVs), FP64 scipy-RK45 twin"
    elif name == "mm_progress_exact":
        g = np.load(GOLDEN)
        lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"], integrator="exact")
        prior = pkg.UniformBox([0, 0, 0], [10, 10, 10], names=["Vmax", "Km", "sigma"])
        n = n_per_gpu or (1 << 20)
        cfg = dict()
        desc = ("CONVERGED MODE, not the reference's likelihood: Michaelis-Menten progress curves in closed form (Wright "
                "omega; the converged solution of the reference's ODE, up to 3e-3 relative from its rtol-1e-3 numbers), FP64")
    elif name == "mm_rate":
        lik = pkg.MMRate.synthetic(10000, precision=32)
        prior = pkg.UniformBox([0, 0, 0], [10, 10, 10], names=["Vmax", "Km", "sigma"])
        n = n_per_gpu or (1 << 22)
        cfg = dict()
        desc = "synthetic Michaelis-Menten rate law, 10k observations, FP32 terms / FP64 accumulation"
    elif name == "mm_rate_suff":
        lik = pkg.MMRate.synthetic(10000, form="sufficient", km_range=(0.0, 10.0))
        prior = pkg.UniformBox([0, 0, 0], [10, 10, 10], names=["Vmax", "Km", "sigma"])
        n = n_per_gpu or (1 << 23)
        cfg = dict()
        desc = ("synthetic Michaelis-Menten rate law, 10k observations, sufficient-statistic form (sum v^2, A(Km), B(Km) "
                "tabulated once; FP64), BASELINE config 4: 2^26 particles over 8 GPUs = 2^23 per GPU")
    elif name == "kinetic":
        cond, base, obs, low, high = kf["cond"], kf["base4"], kf["obs4"], kf["low4"], kf["high4"]
        lik = pkg.KineticRK(cond, obs, base, kf["est4"], n_steps=50)
        prior = pkg.UniformBox(low, high, names=["Af", "Eaf", "Ar", "Ear", "sigma"])
        n = n_per_gpu or (1 << 18)
        cfg = dict()
        desc = "methanation-style plug-flow reactor, d=5, 30 conditions, RK4 x 50 steps, FP64"
    elif name == "kinetic_dae":
        g = np.load(os.path.join(ROOT, "tests", "golden", "dae_synth.npz"))
        lik = pkg.KineticDAE(g["cond"], g["obs"], g["base4"], g["est4"])
        prior = pkg.UniformBox(kf["low4"], kf["high4"], names=["Af", "Eaf", "Ar", "Ear", "sigma"])
        n = n_per_gpu or 1000
        cfg = dict()
        desc = ("the reference's transient fixed-bed reactor (357-unknown DAE, 30 conditions, start-up to 75 s), "
                "implicit Euler + modified Newton, FP64")
    else:
        cond, base, obs = kf["cond"], kf["base16"], kf["obs16"]
        est = np.arange(32, dtype=np.int32)
        lo, hi = base[:32] * 0.8, base[:32] * 1.2
        lik = pkg.KineticRK(cond, obs, base, est, n_steps=50)
        prior = pkg.UniformBox(np.minimum(lo, hi), np.maximum(lo, hi))
        n = n_per_gpu or (1 << 21)
        cfg = dict(fused_sweeps=10, mhstep_num=10, ad_mhstep_num=10, early_exit=False)
        desc = "32-parameter kinetic family (4 LH channels), 30 conditions, RK4 x 50, 10 fused MH sweeps/stage"
    return lik, prior, n, cfg, desc


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU baseline
def _cpu_chunk(args):
    from oracle import mm
    P, t, Pobs, S0 = args
    return mm.sweep_progress(P, t, Pobs, S0, which="scipy")


def cpu_sweep(pool, cores, P, data, per_core=4):
    chunks = [c for c in np.array_split(P, cores * per_core) if len(c)]
    return np.concatenate(pool.map(_cpu_chunk, [(c, *data) for c in chunks]))


def cpu_baseline_mm_progress(target_s=15.0):
    """The reference's likelihood (scipy RK45 through the oracle) on all host cores, one chunk of
    particles per task, over a strided sample of every particle cloud the reference's own N=1000
    run evaluated (prior cloud ... posterior cloud: the golden fixture's 34 sweeps)."""
    import multiprocessing as mp
    g = np.load(GOLDEN)
    data = (g["data_t"], g["data_P"], g["data_S0"])
    cores = os.cpu_count() or 1
    allp = g["sweeps_in"].reshape(-1, 3)
    stride = max(1, int(math.ceil(len(allp) / (cores * 110.0 * target_s))))
    P = allp[::stride]
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_sweep(pool, cores, P[: cores * 4], data)          # warm the workers
        t0 = time.perf_counter()
        cpu_sweep(pool, cores, P, data)
        dt = time.perf_counter() - t0
    return {"value": len(P) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(P)} of the 34000 particle evaluations of the reference's own N=1000 run "
                      f"(every {stride}th, all 34 sweeps), scipy solve_ivp RK45 via oracle.mm, {dt:.1f} s"}


def cpu_config1_run():
    """BASELINE config 1 on the CPU: the reference's own run (N=1000, its seed, its random stream) through the
    oracle's restatement of the loop, likelihood sweeps fanned out over all host cores.  Returns seconds."""
    import multiprocessing as mp
    from oracle import smc
    g = np.load(GOLDEN)
    data = (g["data_t"], g["data_P"], g["data_S0"])
    cores = os.cpu_count() or 1
    st = smc.ReferenceStream(int(g["seed"]))
    p0 = st.prior_uniform([0, 0, 0], [10, 10, 10], 1000)
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_sweep(pool, cores, p0[: cores * 4], data)          # warm the workers
        t0 = time.perf_counter()
        p, lk, tr = smc.run(lambda th: cpu_sweep(pool, cores, th, data), p0, np.zeros(3), np.full(3, 10.0),
                            smc.Settings(), st)
        dt = time.perf_counter() - t0
    return {"seconds": dt, "stages": len(tr.gamma), "sweeps": int(sum(tr.n_mh)) + 1, "cores": cores,
            "log_evidence": tr.log_evidence[-1], "posterior_mean": [float(x) for x in p.mean(0)]}


def _cpu_kin_chunk(args):
    from oracle import kinetic
    th, cond, obs, base, est, n_steps = args
    return kinetic.loglik(th, cond, obs, base, est, n_steps)


def cpu_baseline_kinetic(target_n=65536):
    """The kinetic reactor likelihood on the host cores: the oracle's NumPy twin of the device model (the
    reference's own IDA model cannot be run here), prior-box particles, one chunk per worker."""
    import multiprocessing as mp
    kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
    cond, base, obs, est = kf["cond"], kf["base4"], kf["obs4"], kf["est4"]
    low, high = kf["low4"], kf["high4"]
    cores = os.cpu_count() or 1
    th = np.random.RandomState(1).uniform(low, high, (target_n, len(low)))
    chunks = [c for c in np.array_split(th, cores) if len(c)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_kin_chunk, [(c[:8], cond, obs, base, est, 50) for c in chunks])      # warm the workers
        t0 = time.perf_counter()
        pool.map(_cpu_kin_chunk, [(c, cond, obs, base, est, 50) for c in chunks])
        dt = time.perf_counter() - t0
    return {"value": target_n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{target_n} prior-box particles, d=5, 30 conditions x RK4 x 50 steps, oracle.kinetic (NumPy, "
                      f"vectorised over the particles of a chunk), multiprocessing over all cores, {dt:.1f} s"}


def _cpu_dae_chunk(args):
    from oracle import methanation_dae as dae
    th, cond, obs, base, est = args
    return dae.loglik(th, cond, obs, base, est)


def cpu_baseline_dae(per_core=2):
    """The transient reactor likelihood on the host cores: oracle/methanation_dae.py (NumPy residual, banded LAPACK
    solves; the reference's assimulo / IDA is not installed), particles around the data-generating parameters."""
    import multiprocessing as mp
    g = np.load(os.path.join(ROOT, "tests", "golden", "dae_synth.npz"))
    cond, base, obs, est = g["cond"], g["base4"], g["obs"], g["est4"]
    cores = os.cpu_count() or 1
    n = per_core * cores
    th = base[est] * np.random.RandomState(1).uniform(0.8, 1.25, (n, len(est)))
    chunks = [c for c in np.array_split(th, cores) if len(c)]
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_dae_chunk, [(c, cond, obs, base, est) for c in chunks])
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} particles x 30 conditions (={n * 30} marches), oracle.methanation_dae (same implicit-Euler "
                      f"grid and Newton policy as the device kernel), multiprocessing over all cores, {dt:.1f} s"}


REF_PARTICLES = 128      # particles of one reference-arm step (fixed: the same sample on every box)
DATA_LABEL = {"mm_progress_exact": "observations: the reference's own six CSVs; prior particles: synthetic, U[0,10]^3",
              "mm_progress": "observations: the reference's own six CSVs (SMC_example/data/mm_pseudo_data_0..5.csv, 6x40 "
                             "points, carried in tests/golden/mm_reference_run.npz); prior particles: synthetic, U[0,10]^3",
              "mm_rate": "synthetic", "mm_rate_suff": "synthetic", "kinetic": "synthetic", "kinetic32": "synthetic", "kinetic_dae": "synthetic"}


def run_reference_arm(args):
    """The reference's CPU implementation of the hot path (the oracle's restatement of Micmem_SMC_main.py:98-262 with
    scipy RK45 as the likelihood, fanned out over all host cores) on a bounded sample.  One step is what a step of
    the main arm is - one complete tempered-SMC run, prior -> beta = 1 - with REF_PARTICLES particles instead of
    2^20 (the reference evaluates all N particles in every sweep, out-of-box proposals at the old point)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import smc
    g = np.load(GOLDEN)
    data = (g["data_t"], g["data_P"], g["data_S0"])
    cores = os.cpu_count() or 1
    M = REF_PARTICLES
    cfg = smc.Settings(n_particle=M)
    times, evals, sweeps, stages = [], 0, 0, 0
    with mp.get_context("fork").Pool(cores) as pool:
        loglik = lambda th: cpu_sweep(pool, cores, th, data, per_core=1)
        for it in range(args.warmup + args.steps):
            st = smc.ReferenceStream(int(g["seed"]) + it)          # legacy NumPy stream, the reference's draw order
            p0 = st.prior_uniform([0, 0, 0], [10, 10, 10], M)
            t0 = time.perf_counter()
            _, _, tr = smc.run(loglik, p0, np.zeros(3), np.full(3, 10.0), cfg, st)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
                evals += tr.n_eval
                sweeps += int(sum(tr.n_mh)) + 1
                stages += len(tr.gamma)
    total = sum(times)
    val = evals / total
    k = len(times)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / k,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": DATA_LABEL["mm_progress"],
            "config": {"workload": "mm_progress: Michaelis-Menten tempered SMC, 6x40 progress-curve observations "
                                   "(reference CSVs), FP64",
                       "particles_total": M, "step": "one full tempered-SMC run, prior -> beta=1",
                       "evals_per_step": evals / k, "stages_per_step": stages / k, "sweeps_per_step": sweeps / k},
            "time_to_beta1_s": total / k,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{k} complete tempered-SMC runs of {M} prior particles each (the main arm's step "
                                       f"at {M} instead of 2^20 particles; N evaluations per sweep as in the reference); "
                                       "oracle restatement of the loop, scipy solve_ivp RK45 likelihood, "
                                       "multiprocessing over all cores, one chunk of particles per core"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ roofline helpers
def mm_progress_flops(stats_delta, n_particle_evals, n_ex, n_t):
    """Algorithmic FP64 flop of the progress-curve likelihood from the device's own work counters
    (DESIGN.md 'K1 work model'): 84 flop per step attempt (6 RHS x 3, stage sums 35, y_new 11, error
    norm 13, controller 7), 34 per accepted step (dense-output Q), 18 per observation, 30 per solve
    (initial step)."""
    fev, acc, rej = (int(x) for x in stats_delta[:3])
    return 84.0 * (acc + rej) + 34.0 * acc + 18.0 * n_particle_evals * n_ex * n_t + 30.0 * n_particle_evals * n_ex


def kinetic_flops_per_eval(n_cond, n_steps, M):
    """Algorithmic FP64 flop of one plug-flow reactor likelihood (DESIGN.md K1'), reference formulation, every
    add / multiply / divide / sqrt / exp counted as ONE flop: a right-hand side costs 55 + 31*M (state 20, rate law
    9 + 31 per Langmuir-Hinshelwood channel incl. its 4 Arrhenius factors, density 18, balances 8), a classical RK4
    step 4 right-hand sides + 34, a condition n_steps steps + 75 (outlet flows, residuals)."""
    return float(n_cond) * (n_steps * (4.0 * (55 + 31 * M) + 34.0) + 75.0)


DAE_FLOPS_PER_MARCH = 35 * (51 * 21 * 150 + 51 * 3500 + 4 * (51 * 150 + 51 * 200))
"""Nominal FP64 flop of one transient-reactor march (DESIGN.md K1''): 35 grid steps, each one finite-difference
Jacobian (51 nodes x 21 perturbed node residuals x ~150 flop), one block-tridiagonal factorisation (51 x ~3500) and
~4 modified-Newton iterations (residual 51 x 150 + substitution 51 x 200).  Retries and the early steady-state
exit are not counted: this is a nominal figure, labelled so in the line."""


def resample_microbench(pkg, eng, torch, flush):
    """HBM leg: the whole of K3 (weights -> counts -> offsets -> ancestors -> gather) as the single kernel the
    single-GPU engine runs, on the weights of a real first stage (prior cloud, the tempering step's own max / gm /
    sum_w), inputs flushed from L2.  Algorithmic bytes per particle: 8 (weight) + 4 + 4 (ancestor
    write / read) + 2*(d+1)*8 (state read + write)."""
    n, D1 = eng.n, eng.d + 1
    eng.sample_prior()
    eng.sim_particle()
    t = eng.temper(0.0)
    ms = []
    for _ in range(5):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._ck(eng.lib.smcb_resample_fused(eng.h, eng.lk.data_ptr(), None, n, n, 0, 0, n, eng.scal.data_ptr(), t["gm"],
                                            eng.scal[1:].data_ptr(), 0.375, eng.state.data_ptr(), n, D1,
                                            eng.state2.data_ptr(), n, eng.anc.data_ptr(), None,
                                            eng.icnt[6:7].data_ptr(), eng._stream))
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = float(np.median(ms))
    nbytes = n * (8 + 4 + 4 + 2 * D1 * 8)
    return best, nbytes


# ------------------------------------------------------------------------------------ main arm
def main():
    args = parse()
    if args.cpu_baseline_only:
        if args.workload == "kinetic":
            print(json.dumps(cpu_baseline_kinetic()), flush=True)
            return
        if args.workload == "kinetic_dae":
            print(json.dumps(cpu_baseline_dae()), flush=True)
            return
        out = cpu_baseline_mm_progress()
        out["config1_run"] = cpu_config1_run()
        print(json.dumps(out), flush=True)
        return
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import smcb200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sampler path has no CPU fallback")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the first communicator is created; stdout carries exactly one
        # JSON line here, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            comm = pkg.NcclComm.from_torch_distributed()     # NCCL inside libsmcb200.so; torch only carries the id
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n_gpus = world

    if args.total_particles:
        assert args.total_particles % world == 0, "--total-particles must be a multiple of the number of GPUs"
        args.particles = args.total_particles // world
    lik, prior, n_loc, cfg_kw, desc = make_workload(pkg, args.workload, args.particles)
    N = n_loc * world
    cfg = pkg.Settings(n_particle=N, **cfg_kw)
    eng = pkg.Engine(lik, prior, cfg, comm=comm)
    eng.sample_prior()
    prior_dev = eng.state[: eng.d].clone()            # prior particles resident in HBM
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def flush():
        flush_buf.fill_(1)                            # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.state[: eng.d].copy_(prior_dev)
        return eng.run()

    for _ in range(args.warmup):
        flush()
        res = step()
    stats0 = eng.loglik_stats() if args.workload == "mm_progress" else None
    if args.workload == "mm_progress":
        eng.kernel_profile(True)          # CUDA events around the bulk and the tail kernel, recorded by the library
    launches0 = eng.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total, evals, decided, stages, sweeps = 0.0, 0, 0, 0, 0
    barrier()
    for _ in range(args.steps):
        flush()                                       # L2 flushed between timed iterations (untimed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = step()
        e1.record()
        e1.synchronize()
        barrier()
        ms_total += e0.elapsed_time(e1)
        evals += res.n_eval - res.n_eval_cut      # evaluations carried through every observation
        decided += res.n_eval                     # + proposals whose rejection was proven earlier
        stages += len(res.betas)
        sweeps += sum(res.n_mh)
        assert res.reached_one, "tempering did not reach beta = 1"
    clk = clocks.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    kprof = eng.kernel_profile() if args.workload == "mm_progress" else None
    if kprof is not None:
        eng.kernel_profile(False)
    st1 = eng.loglik_stats() if args.workload == "mm_progress" else None
    # one more, UNTIMED step with an event pair around every kernel group of the host loop: the per-group
    # breakdown (kernel_ms, sweep_group_ms).  Recording ~500 event pairs costs the host ~8 ms per run, which is
    # why it is kept out of the timed steps.
    eng.enable_profiling(True)
    flush()
    step()
    prof = eng.profile_summary()
    lik_ms = eng.profile_events("loglik")
    eng.enable_profiling(False)
    prof_steps = 1
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    secs = ms_total * 1e-3
    value = evals / secs

    # ---- roofline of the dominant kernel (rank 0's shard) ---------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    fma = np.zeros(2)
    eng._ck(eng.lib.smcb_measure_fma_peak(eng.h, fma.ctypes.data))
    n_lik, ms_lik = prof.get("loglik", (0, 0.0))
    if "mh_fused" in prof:
        n_lik, ms_lik = n_lik + prof["mh_fused"][0], ms_lik + prof["mh_fused"][1]
    # scale the one profiled step to the timed region so that shares are taken against the same total
    n_lik, ms_lik = n_lik * args.steps, ms_lik * args.steps
    roofline = {"bound": "fp64", "kernel": "likelihood", "achieved": None, "peak": fma[0] / 1e12, "unit": "TFLOP/s",
                "frac": None, "traffic": None,
                "peak_source": "FP64 FMA micro-benchmark run in this process (smcb_measure_fma_peak); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "share_of_step": ms_lik / ms_total if ms_total else None,
                "avg_launch_ms": ms_lik / max(n_lik, 1), "launches": n_lik}
    roofline_tail = None
    if args.workload == "mm_progress":
        d_stats = st1[4:8] - stats0[4:8]
        tail_attempts = int(st1[15] - stats0[15])
        local_evals = evals // world
        n_ex, n_t = lik.t.shape
        flops_all = mm_progress_flops(d_stats, local_evals, n_ex, n_t)
        per = len(lik_ms)
        bulk_ms, tail_ms, n_sw = kprof
        # the bulk kernel's share of the algorithmic work: every step that was not taken inside the tail kernel
        att_all = float(d_stats[1] + d_stats[2])
        flops_bulk = flops_all * (1.0 - tail_attempts / max(att_all, 1.0))
        roofline.update(
            kernel=f"mm_bulk_kernel (one launch per likelihood sweep; all solves of up to {eng.mm_budget} attempted steps)",
            achieved=flops_bulk / (bulk_ms * 1e-3) / 1e12, launches=n_sw, avg_launch_ms=bulk_ms / max(n_sw, 1),
            share_of_step=bulk_ms / ms_total if ms_total else None,
            sweep_group_ms_last_step=[round(x, 3) for x in lik_ms[-per:]],
            sweep_group_share_of_step=ms_lik / ms_total if ms_total else None,
            rhs_evals_per_particle_eval=float(d_stats[0]) / max(local_evals, 1),
            rejected_step_fraction=float(d_stats[2]) / max(att_all, 1.0),
            flop_model="84/attempted step + 34/accepted step + 18/observation + 30/solve (DESIGN.md K1), step counts "
                       "from the device counters; proposals rejected early are not counted",
            # DRAM bytes of one launch from the ncu --set full capture of a posterior-cloud sweep at 2^20 particles
            # (profiles/ncu_full_r02_bulk.md: 60.7 MB read + 41.4 MB written), scaled to this shard's
            # particle count - not re-measured here; the kernel moves 61 B per particle and is nowhere near HBM-bound
            traffic=(102.12e6 / (1 << 20)) * eng.n,
            traffic_source="ncu --set full capture profiles/ncu_full_r02_bulk.md (a posterior-cloud sweep at 2^20 particles), per "
                           "particle x particles of a sweep; the kernel moves ~100 B per particle and is nowhere near HBM-bound")
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        longest = int(st1[16])
        roofline_tail = {
            "bound": "latency", "kernel": "mm_tail_kernel (deferred solves, one per lane; duration = longest solve)",
            "launches": n_sw, "ms_per_step": tail_ms / args.steps, "share_of_step": tail_ms / ms_total if ms_total else None,
            "attempted_steps_in_tail_per_step": tail_attempts / args.steps,
            "first_sweep_group_ms": round(lik_ms[0], 3) if lik_ms else None,
            "longest_solve": {"attempts": longest >> 32, "cycles_per_attempt": longest & 0xffffffff},
            "floor_cycles_per_attempt": 456,
            "note": "a solve is a strictly serial chain: per attempted step 6 x (MUFU.RCP64H + 3 dependent FMAs) and the "
                    "step-size controller (2 MUFU + ~12 dependent FP64 operations), ~430 cycles of dependent latency; 456 "
                    "cycles is the same step measured with the solve alone on the GPU (profiles/ncu_full_r02_tail.md); "
                    "round 1's spelling took 707"}
    if args.workload == "mm_rate":
        # SURVEY 8(d) form A: 6 flop per (particle, observation) with the reciprocal counted as one; the kernel
        # spends 5.5 FMA-pipe lane slots + 0.5 MUFU per term, so lane-slot occupancy is the tighter figure
        terms = float(evals // world) * lik.n_obs
        lane_peak = fma[1] / 2.0                      # FMA-pipe lane slots per second (an FFMA is 2 flop)
        roofline.update(bound="fp32", kernel="mm_rate_kernel_f32 (4 particles per thread, paired reciprocals, f32x2)",
                        achieved=6.0 * terms / (ms_lik * 1e-3) / 1e12, peak=fma[1] / 1e12,
                        peak_source="FP32 FFMA micro-benchmark run in this process (smcb_measure_fma_peak)",
                        fma_pipe_lane_slot_frac=5.5 * terms / (ms_lik * 1e-3) / lane_peak,
                        terms_per_s=terms / (ms_lik * 1e-3))
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
    if args.workload == "mm_progress_exact":
        # per observation: Taylor predictor 14 + one order-4 correction (1 log, 2 divisions, 12 others) + residual 3
        fl = 32.0 * lik.n_obs
        ev = float(evals // world)
        roofline.update(kernel="mm_exact_kernel (one thread per particle, Wright omega by continuation in time)",
                        achieved=fl * ev / (ms_lik * 1e-3) / 1e12, flops_per_eval=fl,
                        flop_model="32 flop per observation (predictor 14, one Fritsch-Shafer-Crowley correction 15 with "
                                   "log and divisions counted once, residual 3)")
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
    if args.workload == "mm_rate_suff":
        # HBM-bound: 24 B of parameters + 1 B of mask in, 8 B out per evaluated particle
        nbytes = 33.0 * float(evals // world)
        roofline.update(bound="hbm", kernel="mm_rate_kernel_suff (two 14-term Clenshaw sums per particle)",
                        achieved=nbytes / (ms_lik * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s", peak_source=hbm_src,
                        bytes_model="33 B per evaluated particle (3 parameters + mask in, log-likelihood out)")
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
    if args.workload in ("kinetic", "kinetic32", "kinetic_dae"):
        evals_step = float(evals) / args.steps / world            # completed evaluations of one step on this rank
        ms_step = ms_lik / args.steps                             # likelihood launch groups of the one profiled step
        if args.workload == "kinetic_dae":
            fl = DAE_FLOPS_PER_MARCH * lik.cond.shape[0]
            roofline.update(kernel="dae_march_kernel (one block per (particle, condition))", bound="latency",
                            flop_model="NOMINAL 35 steps x (FD Jacobian + block-tridiagonal factorisation + 4 Newton "
                                       "iterations) per march (bench.py DAE_FLOPS_PER_MARCH); the kernel is bound by the "
                                       "serial chain over the 51 nodes and its barriers, not by arithmetic")
        else:
            M = lik.n_pairs // 4
            fl = kinetic_flops_per_eval(lik.cond.shape[0], lik.n_steps, M)
            roofline.update(kernel=f"kinetic_ssr_kernel<M={M}> (one thread per (particle, condition), RK4 x {lik.n_steps})",
                            flop_model=f"{fl:.0f} flop per likelihood = n_cond x (n_steps x (4 x (55 + 31 M) + 34) + 75), "
                                       "every add/mul/div/sqrt/exp counted once (the kernel spends 4 / 7 / 10 FP64 "
                                       "operations on a division / root / exponential)")
        roofline.update(achieved=fl * evals_step / (ms_step * 1e-3) / 1e12, flops_per_eval=fl,
                        evals_per_step_this_rank=evals_step)
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
    roofline_hbm = None
    if world == 1:
        g_ms, g_bytes = resample_microbench(pkg, eng, torch, flush)
        roofline_hbm = {"bound": "hbm", "kernel": "resample_fused_kernel (all of K3 in one launch: weights, counts, look-back "
                                                  "scan, ancestors, gather of the particle state)",
                        "achieved": g_bytes / (g_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": g_bytes / (g_ms * 1e-3) / 1e9 / hbm_peak,
                        # DRAM bytes of one launch from the ncu --set full capture (profiles/ncu_full_r02_resample.md:
                        # 33.6 MB read + 1.0 MB written at 2^20 particles, d = 3: inside a run the 32 MiB state and its
                        # copy live in the 126 MB L2), scaled to this launch's particle count - not re-measured here
                        "traffic": (34.6e6 / (1 << 20)) * eng.n if eng.d == 3 else None,
                        "traffic_source": "ncu capture profiles/ncu_full_r02_resample.md (L2-resident state), per particle",
                        "bytes_model": "8 (weight) + 4 + 4 (ancestor) + 2 (d+1) 8 (state) per particle, SURVEY.md 8(d)",
                        "peak_source": hbm_src, "launch_ms": g_ms, "bytes": g_bytes}

    # ---- BASELINE config 1: the reference's own problem size (N = 1000, its seed and random stream) -----------
    config1 = None
    if args.workload == "mm_progress" and world == 1:
        from importlib import import_module
        LegacyNumpyStream = import_module(pkg.__name__ + ".reference_api").LegacyNumpyStream
        g1 = np.load(GOLDEN)
        eng1 = pkg.Engine(lik, prior, pkg.Settings(n_particle=1000, scan_mode="sequential"))
        t1 = []
        for it in range(3):
            st_ref = LegacyNumpyStream(int(g1["seed"]))
            p0 = st_ref.prior_uniform([0, 0, 0], [10, 10, 10], 1000)
            r1 = eng1.run(p0, stream=st_ref)
            t1.append(r1.seconds)
        config1 = {"workload": "reference MM run: N=1000, seed 20250205, NumPy legacy stream (same draws as the reference)",
                   "time_to_beta1_s": float(np.median(t1[1:])), "stages": len(r1.betas), "sweeps": int(sum(r1.n_mh)) + 1,
                   "log_evidence": r1.log_evidence, "posterior_mean": [float(x) for x in r1.particles.mean(0)]}
        eng1.close()

    # ---- end to end through the public API, host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        host_p = torch.empty((n_loc, eng.d), dtype=torch.float64).pin_memory()
        host_p.copy_(prior_dev.t())
        eng.close()
        del eng
        torch.cuda.empty_cache()
        e_evals, e_t = 0, 0.0
        E_WARM, e_n = 2, max(1, min(args.steps, 3))    # untimed warm-up calls (pool, pinned staging), timed calls
        for it in range(E_WARM + e_n):
            barrier()
            t0 = time.perf_counter()
            r = pkg.run(lik, prior, particles=host_p, settings=cfg, comm=comm)     # H2D + run inside
            post, post_lk = r.particles, r.lk                                      # D2H of the result
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                import torch.distributed as dist
                tt = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local))
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            print(f"e2e call {it}: {dt * 1e3:.1f} ms (device {r.seconds * 1e3:.1f} ms)", file=sys.stderr, flush=True)
            if it >= E_WARM:
                e_evals += r.n_eval - r.n_eval_cut
                e_t += dt
        e2e = {"value": e_evals / e_t, "unit": UNIT, "h2d_bytes_per_step": int(N * prior.d * 8),
               "d2h_bytes_per_step": int(N * (prior.d + 1) * 8), "seconds_per_step": e_t / e_n,
               "api": "smcb200.run(likelihood, prior, pinned host particles, settings) -> Result (host arrays)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ("mm_progress", "kinetic", "kinetic_dae"):
        # separate process: the worker pool forks, which must not happen under a live CUDA context
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only", "--workload",
                              args.workload], capture_output=True, text=True, timeout=600)
        cpu = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else {"error": out.stderr[-300:]}
        if args.workload in ("kinetic", "kinetic_dae") and "value" in cpu:
            # the CPU cannot run 1e7 evaluations in the bench's time budget: the figure below is the measured CPU
            # likelihood rate applied to the evaluations the GPU run needed - an extrapolation, labelled as such
            cpu["time_to_beta1_s_extrapolated"] = (evals / args.steps) / cpu["value"]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.total_particles else "weak", "vs_baseline": None,
                "dtype": "f64" if args.workload != "mm_rate" else "f32",
                "data": DATA_LABEL[args.workload],
                "config": {"workload": f"{args.workload}: {desc}", "particles_total": N, "particles_per_gpu": n_loc,
                           "d": prior.d, "observations": int(lik.n_obs), "temper_rule": cfg.temper_rule,
                           "scan_mode": cfg.scan_mode, "l2": "256 MiB flush between timed steps",
                           "step": "one full tempered-SMC run, prior -> beta=1"},
                "time_to_beta1_s": secs / args.steps, "stages_per_step": stages / args.steps,
                "mh_sweeps_per_step": sweeps / args.steps, "likelihood_sweeps_timed": sweeps + args.steps,
                "evals_per_step": evals / args.steps,
                "evals_note": "value counts likelihood evaluations integrated over every observation (first sweep + "
                              "in-box MH proposals); proposals whose rejection was proven before the last observation "
                              "(exact early rejection) are NOT counted and are reported in decisions_per_step; the "
                              "reference would evaluate reference_evals_per_step (N per sweep, out-of-box ones too)",
                "decisions_per_step": decided / args.steps,
                "reference_evals_per_step": res.n_eval_reference,
                "log_evidence": res.log_evidence, "posterior_mean": [float(x) for x in res.particles.mean(0)],
                "roofline": roofline, "roofline_tail": roofline_tail, "roofline_hbm": roofline_hbm,
                "kernel_ms_per_step": {k: {"groups": v[0], "ms": v[1]} for k, v in prof.items()},
                "kernel_ms_note": "device time of each kernel group in one extra, untimed step",
                "fp32_fma_peak_tflops": fma[1] / 1e12,
                "config1": config1, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
