#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract in the task statement, tier section 4).

A "step" is ONE fused evaluation of fg! (composite -> Poisson logL -> gradient; src/fitting/solvers.jl:20-38)
over the BASELINE.json headline stack: 200x300 bins x 2400 templates (60 ages x 40 [M/H]), Float64,
Poisson-sampled synthetic Hess diagram.  Each rank holds one such stack (N>1: a bin-row shard of an N x larger
Hess diagram, [logL, G] all-reduced inside the finalize kernel every step => weak scaling).

  value       evaluations/s with everything resident in HBM, timed with CUDA events on the launching stream; ranks aligned
              on the device (two untimed all-reduced steps, then e0 in-stream), max over ranks
  e2e         same through the reference-facing C-ABI call sfh_eval_fg with HOST buffers (H2D coeffs, D2H [-logL, G])
  fg_hier     the call fit_sfh / sample_sfh make: sfh_eval_fg_hier (PowerLawMZR + GaussianDispersion), wall clock per call
  roofline    algorithmic bytes of the fused kernel / its event-timed duration, vs MEASURED_PEAKS.json hbm_gbs
  config5     BASELINE config 5 as STRONG scaling: one 10^6 x 10^4 Float32 stack (40 GB) split over the N GPUs
  parity      N = 1: logL and all 2400 gradient components vs the oracle; N > 1: ranks bit-identical and the all-reduced
              answer == the WHOLE stack evaluated on rank 0's GPU alone (config 3 x N and config 5).  The run fails otherwise.
  cpu_baseline  the faster of the oracle's two threaded two-pass ports of the reference algorithm (BLAS gemv route / OpenMP nest)
                on this box's host cores (N = 1 only)

  --impl reference : the reference arm = that same CPU port, all host threads (set here, not inherited from the launcher),
                     on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time


def _host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# The CPU legs use every host core.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently turn the
# reference arm into a one-core run (round 1's N>1 "vs_reference" ratios): the thread counts are set here, before numpy
# (OpenBLAS) and the OpenMP oracle are loaded, and re-asserted at run time through threadpoolctl / omp_set_num_threads.
if "--impl" in sys.argv and "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_host_cores())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB, NJ, NK = 200 * 300, 60, 40
NT = NJ * NK
SEED = 94823
METRIC = "fused loglikelihood+gradient evaluations/sec (fg!), 200x300 bins x 2400 templates Float64"
UNIT = "evals/s"
WORKLOAD = "config3: fit_sfh PowerLawMZR+GaussianDispersion stack, 60000 bins x 2400 templates (60 logAge x 40 [M/H]), Float64, Poisson data"


def truth_coeffs():
    """r_jk of mzr_test.jl:52-63 scaled to config 3: R = 1e6 U(0,1), PowerLawMZR(1,-2,6), GaussianDispersion(0.2).
    Host float64 formulae (one-shot set-up, not the timed path)."""
    rng = np.random.Generator(np.random.Philox(SEED))
    uA = np.linspace(10.1, 6.6, NJ)
    uM = np.linspace(-2.5, 0.0, NK)
    R = rng.random(NJ) * 1e6
    cum = np.cumsum(R)                      # ages already sorted oldest -> youngest
    mu = -2.0 + 1.0 * (np.log10(cum) - 6.0)
    A = np.exp(-(((uM[None, :] - mu[:, None]) / 0.2) ** 2) / 2)
    r = (A * R[:, None] / A.sum(axis=1, keepdims=True)).reshape(-1)
    return r


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling guide recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key="fused_kernel_dram_bytes_per_launch"):
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


def host_stack(nb, nt, x, seed=SEED):
    """Synthetic stack on the HOST for the CPU arms (numpy Philox): U(0,1)/1e5 templates, Poisson(M x) data."""
    rng = np.random.Generator(np.random.Philox(seed))
    M = np.empty((nb, nt), dtype=np.float64, order="F")
    for j0 in range(0, nt, 100):
        j1 = min(nt, j0 + 100)
        M[:, j0:j1] = rng.random((nb, j1 - j0)) / 1e5
    data = rng.poisson(M @ x).astype(np.float64)
    return M, data


def cpu_routes(M, data, x):
    """The two CPU restatements of the reference's flat fg!, both on all host threads:
    'blas'   -- the route Julia takes: gemv 'N' / Poisson + residual loops / gemv 'T' through OpenBLAS (oracle.fg_blas);
    'openmp' -- the oracle's own two-pass loop nest, no BLAS (oracle.fg_omp).  The FASTER one is the reported baseline."""
    import oracle as O
    G = np.empty(M.shape[1]); Cm = np.empty(M.shape[0])
    return {"blas": (lambda: O.fg_blas(x, M, data), O.blas_threads()),
            "openmp": (lambda: O.fg_omp(x, M, data, G=G, Cm=Cm), O.num_threads())}


def time_cpu(M, data, x, budget_s=12.0, min_steps=3, max_steps=200):
    """cpu_baseline leg: median evaluation time of each route inside half the budget; returns the faster one."""
    out = {}
    for name, (fn, threads) in cpu_routes(M, data, x).items():
        fn()                                    # warm-up (page-in)
        ts = []
        t_end = time.perf_counter() + budget_s / 2
        while (len(ts) < min_steps or time.perf_counter() < t_end) and len(ts) < max_steps:
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        out[name] = (float(np.median(ts)), len(ts), threads)
    best = min(out, key=lambda k: out[k][0])
    return best, out


def time_cpu_one_thread(M, data, x, reps=3):
    """The BLAS route with OpenBLAS held to ONE thread: the setting of the reference's own benchmark suite
    (benchmark/benchmarks.jl:5, BLAS.set_num_threads(1)) and of tsample_sfh (generic_fitting.jl:593).  evals/s, or None."""
    try:
        from threadpoolctl import threadpool_limits
        import oracle as O
        with threadpool_limits(limits=1, user_api="blas"):
            O.fg_blas(x, M, data)
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                O.fg_blas(x, M, data)
                ts.append(time.perf_counter() - t0)
        return 1.0 / float(np.median(ts))
    except Exception:
        return None


CPU_ROUTE_TEXT = {"blas": "gemv 'N' / Poisson + residual loops / gemv 'T' through numpy's OpenBLAS, the route Julia's mul! takes",
                  "openmp": "the oracle's OpenMP two-pass loop nest, no BLAS"}


def run_reference(args, real_stdout):
    """--impl reference: the reference's algorithm (two-pass gemv 'N' / Poisson / residual / gemv 'T') on the host
    cores.  Julia cannot run in this image, so this is the oracle PORT (cpu_baseline.kind = "port"): both restatements are
    warmed up and given a short trial; the faster one runs the timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle as O
    ncores = _host_cores()
    O.set_num_threads(ncores)                       # OpenMP nest
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncores, user_api="blas")   # numpy's OpenBLAS
    except Exception:
        pass
    x = truth_coeffs() * 1.02
    M, data = host_stack(NB, NT, x)
    routes = cpu_routes(M, data, x)
    trial = {}
    for name, (fn, _) in routes.items():
        for _ in range(max(args.warmup, 1)):
            fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        trial[name] = (time.perf_counter() - t0) / 3
    best = min(trial, key=trial.get)
    fn, cores = routes[best]
    per_step = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        fn()
        per_step.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    other = [k for k in routes if k != best][0]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: one full-size stack on rank 0, host cores only"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "route": best,
                             "median_step_value": 1.0 / float(np.median(per_step)),   # shared hosts stall now and then: the mean (value) carries that
                             "blas_one_thread": time_cpu_one_thread(M, data, x),
                             "sample": f"{args.steps} full evaluations of the 60000x2400 F64 stack; port of fitting_base.jl:55-65,84-96,"
                                       f"265-285 (julia not installed): {CPU_ROUTE_TEXT[best]}; the slower restatement "
                                       f"({other}: {CPU_ROUTE_TEXT[other]}) ran at {1.0 / trial[other]:.1f} evals/s in a 3-step trial"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=real_stdout, flush=True)
    return 0


def _protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not break it:
    fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def kernel_name(info, dtype):
    if info.variant == 4:
        vec = 16 // np.dtype(dtype).itemsize
        return "sfh_fg_fused2_kernel<%s,%d,true>" % ("double" if vec == 2 else "float", info.tile_bins // vec)
    return "sfh_fg_fused_kernel<%s,%d,%d,true,false>" % ("double" if np.dtype(dtype).itemsize == 8 else "float",
                                                          info.tile_bins, info.consumer_warps)


def tiling(info):
    return {"variant": info.variant, "tile_bins": info.tile_bins, "cluster": info.cluster, "chunks_per_tile": info.chunks_per_tile,
            "ring_slots": info.ring_slots, "n_clusters": info.n_clusters, "consumer_warps": info.consumer_warps}


class Harness:
    """One process = one GPU.  Everything that touches torch / the library lives here so that the CPU arm never imports them."""

    def __init__(self, args):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import sfh_b200 as S
        self.C, self.torch, self.dist, self.S, self.L = C, torch, dist, S, S._lib
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream()     # an explicit non-default stream: the kernels AND the timing events live on it
        torch.cuda.set_stream(self.stream)
        self.dp = C.POINTER(C.c_double)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def new_ctx(self, ds):
        ctx = ds.new_ctx(self.stream.cuda_stream)
        if self.world > 1 and os.environ.get("SFH_BENCH_NO_EXCHANGE") != "1":
            self.S.init_library_comm(ctx)     # NCCL communicator + (unless SFH_NO_P2P=1) the fused one-shot NVLink all-reduce
        return ctx

    def gather(self, v):
        """[v on rank 0, v on rank 1, ...] on every rank."""
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world == 1:
            return [float(v)]
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def comm_mode(self, ctx):
        nr, rk, mode = self.C.c_int(), self.C.c_int(), self.C.c_int()
        self.L.check(self.L.lib.sfh_ctx_comm_info(ctx.handle, self.C.byref(nr), self.C.byref(rk), self.C.byref(mode)))
        return {0: "single GPU", 1: "NCCL all-reduce on the kernel's stream",
                2: "one-shot NVLink peer-memory all-reduce fused into the finalize kernel"}[mode.value]

    def timed_device_loop(self, ctx, d_x, d_out, steps, warmup, sampler=None):
        """K device-resident evaluations timed with CUDA events on the launch stream.  Ranks are aligned ON THE DEVICE: two
        untimed all-reduced steps are enqueued after the host barrier and e0 is recorded right behind them in-stream -- an
        all-reduced step cannot finish before every rank has contributed, so all ranks pass e0 within an NVLink latency of
        each other and the first timed step does not absorb the host barrier's exit skew."""
        L, C, torch = self.L, self.C, self.torch
        def step():
            L.check(L.lib.sfh_enqueue_fg(ctx.handle, d_x.data_ptr(), d_out.data_ptr(), 1))
        for _ in range(warmup):
            step()
        self.barrier()
        st0, st1 = L.sfh_stats(), L.sfh_stats()
        if sampler:
            sampler.start()
        step(); step()
        L.check(L.lib.sfh_ctx_stats(ctx.handle, C.byref(st0)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            step()
        e1.record(self.stream)
        L.check(L.lib.sfh_ctx_stats(ctx.handle, C.byref(st1)))
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        return ms, int(st1.kernel_launches - st0.kernel_launches)

    def parity_sharded_vs_whole(self, make_whole, x, d_out, nt):
        """N > 1: (i) every rank holds bit-identical [logL, G]; (ii) rank 0 also builds the WHOLE stack (all shards' rows) on its
        own GPU and evaluates it alone: the all-reduced sharded answer must match it (logL 1e-12; every gradient component
        1e-10 relative, plus 1e-12 of the largest component for those that nearly cancel)."""
        torch, dist = self.torch, self.dist
        mine = d_out.view(torch.int64).clone()
        allv = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine)
        identical = all(bool(torch.equal(v, allv[0])) for v in allv)
        res = {"ranks_bit_identical": identical}
        if self.rank == 0:
            whole = make_whole()
            cw = whole.new_ctx(self.stream.cuda_stream)
            d_x = torch.tensor(x, dtype=torch.float64, device="cuda")
            d_w = torch.zeros(1 + nt, dtype=torch.float64, device="cuda")
            self.L.check(self.L.lib.sfh_enqueue_fg(cw.handle, d_x.data_ptr(), d_w.data_ptr(), 1))
            torch.cuda.synchronize()
            w, sh = d_w.cpu().numpy(), d_out.cpu().numpy()
            res["logl_rel"] = float(abs(sh[0] - w[0]) / abs(w[0]))
            tol = 1e-10 * np.abs(w[1:]) + 1e-12 * np.abs(w[1:]).max()
            res["grad_max_err_over_tol"] = float(np.max(np.abs(sh[1:] - w[1:]) / tol))
            res["whole_stack_tiling"] = tiling(whole.info())
            cw.close(); whole.close()
            del cw, whole
            torch.cuda.empty_cache()
        self.barrier()
        flag = torch.tensor([1.0 if (identical and res.get("logl_rel", 0.0) <= 1e-12 and res.get("grad_max_err_over_tol", 0.0) <= 1.0) else 0.0],
                            dtype=torch.float64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["ok"] = bool(flag.item() == 1.0)
        return res


def run_config5(H, args, peak):
    """BASELINE.json config 5, STRONG scaling: ONE 10^6-bin x 10^4-template Float32 stack (40 GB), bin rows split evenly over the
    N GPUs (generated per shard on the device from a counter-based stream, so every N sees the same matrix), [logL, G]
    (10001 doubles) all-reduced every evaluation."""
    S, L, C, torch = H.S, H.L, H.C, H.torch
    nb5, nt5, seed5 = 1000 * 1000, 10000, 58392
    rng = np.random.Generator(np.random.Philox(seed5))
    x5 = rng.random(nt5)
    b, e = S.shard_rows(nb5, H.world, H.rank, align=128)
    ds = S.DeviceStack.synthetic(nb5, nt5, np.float32, seed=seed5, scale=1.0, x_true=x5, device=H.local, rows=(b, e),
                                 tile_bins=args.tile, cluster=args.cluster, consumer_warps=args.nw, variant=args.variant)
    info = ds.info()
    assert info.fused == 1
    ctx = H.new_ctx(ds)
    d_x = torch.tensor(x5 * 1.02, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(1 + nt5, dtype=torch.float64, device="cuda")
    steps = max(3, min(args.steps, 20))
    ms, _ = H.timed_device_loop(ctx, d_x, d_out, steps, 3)
    ms_eval = ms / steps
    ms_e, ms_k = C.c_double(), C.c_double()
    xh = np.ascontiguousarray(x5 * 1.02)
    L.check(L.lib.sfh_time_fg(ctx.handle, xh.ctypes.data_as(H.dp), 5, 1, 0, C.byref(ms_e), C.byref(ms_k)))
    rows = e - b
    bytes_shard = rows * nt5 * 4 + rows * 8 + 2 * nt5 * 8 + 8
    kernel_ms = H.max_over_ranks(ms_k.value)
    out = {"workload": "config5: 10^6 bins x 10^4 templates Float32 (40 GB), bin rows sharded over the GPUs, Poisson data",
           "scaling": "strong", "n_gpus": H.world, "steps": steps, "ms_per_eval": ms_eval, "evals_per_s": 1e3 / ms_eval,
           "aggregate_GBps": (nb5 * nt5 * 4 + nb5 * 8) / (ms_eval * 1e-3) / 1e9,
           "per_gpu": {"rows": rows, "bytes_alg_per_launch": bytes_shard, "kernel_ms": kernel_ms,
                       "roofline_frac_kernel": bytes_shard / (kernel_ms * 1e-3) / 1e9 / peak,
                       "roofline_frac_step": bytes_shard / (ms_eval * 1e-3) / 1e9 / peak, "kernel": kernel_name(info, np.float32)},
           "tiling": tiling(info), "exchange": H.comm_mode(ctx)}
    torch.cuda.synchronize()
    out["neg_logL"] = float(-d_out[0].item())
    if H.world > 1:
        out["parity"] = H.parity_sharded_vs_whole(
            lambda: S.DeviceStack.synthetic(nb5, nt5, np.float32, seed=seed5, scale=1.0, x_true=x5, device=H.local), x5 * 1.02, d_out, nt5)
        assert out["parity"]["ok"], out["parity"]
    ctx.close(); ds.close()
    return out


def run_hier(H, ctx, steps, warmup, peak, bytes_alg):
    """The call fit_sfh / sample_sfh make every iteration: the MZR-hierarchical fg! (mzr.jl:84-215) through the host-synchronous
    C-ABI call sfh_eval_fg_hier (Nj + 3 variables in, [-logL, G] out), wall-clock per call."""
    L, C = H.L, H.C
    uA = np.linspace(10.1, 6.6, NJ); uM = np.linspace(-2.5, 0.0, NK)
    logAge = np.ascontiguousarray(np.repeat(uA, NK)); MH = np.ascontiguousarray(np.tile(uM, NJ))
    rng = np.random.Generator(np.random.Philox(SEED))
    v = np.ascontiguousarray(np.concatenate([rng.random(NJ) * 1e6, [1.0, -2.0, 0.2]]) * 1.02)
    nj = C.c_int64()
    L.check(L.lib.sfh_hier_bind(ctx.handle, logAge.ctypes.data_as(H.dp), MH.ctypes.data_as(H.dp), C.byref(nj)))
    fixed = np.array([6.0, 0, 0, 0]); mask = (C.c_uint8 * 3)(1, 1, 1)
    G = np.empty(NJ + 3); nl = C.c_double()
    fn, h_ctx, p_fixed, p_v, p_nl, p_G = L.lib.sfh_eval_fg_hier, ctx.handle, fixed.ctypes.data_as(H.dp), v.ctypes.data_as(H.dp), C.byref(nl), G.ctypes.data_as(H.dp)
    def call():
        L.check(fn(h_ctx, 0, p_fixed, 0, p_v, mask, p_nl, p_G))
    for _ in range(warmup):
        call()
    H.barrier()
    per_call = np.empty(steps)
    t0 = time.perf_counter()
    for i in range(steps):
        t1 = time.perf_counter()
        call()
        per_call[i] = time.perf_counter() - t1
    wall = H.max_over_ranks(time.perf_counter() - t0)
    ms = 1e3 * wall / steps
    ms_median = H.max_over_ranks(1e3 * float(np.median(per_call)))   # (a 20-call mean over 8 host processes carries any one hiccup)
    assert np.isfinite(nl.value) and np.all(np.isfinite(G))
    return {"what": "sfh_eval_fg_hier (PowerLawMZR + GaussianDispersion, 60 ages + 3 parameters), host buffers, wall clock per call",
            "ms_per_eval": ms, "ms_per_eval_median": ms_median, "evals_per_s": H.world * 1e3 / ms, "h2d_bytes_per_step": (NJ + 3) * 8 + 8, "d2h_bytes_per_step": (NJ + 4) * 16,
            "roofline_frac_end_to_end": bytes_alg / (ms * 1e-3) / 1e9 / peak}


def main():
    real_stdout = _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the strong-scaling 40 GB block")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--nw", type=int, default=0, help="consumer warps per CTA of the cluster-tile kernel (8 or 16; 0 = auto)")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 cluster-tile kernel, 4 warp-specialised stream kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    H = Harness(args)
    C, torch, dist, S, L = H.C, H.torch, H.dist, H.S, H.L
    world, rank, local = H.world, H.rank, H.local
    peak, peak_src = measured_peaks()

    # ---- headline: one config-3 stack per rank, generated on device (counter-based: shards of one N x larger diagram)
    x_true = truth_coeffs()
    nb_total = NB * world
    mk = dict(tile_bins=args.tile, cluster=args.cluster, consumer_warps=args.nw, variant=args.variant)
    ds = S.DeviceStack.synthetic(nb_total, NT, np.float64, seed=SEED, scale=1e-5, x_true=x_true, device=local,
                                 rows=(rank * NB, (rank + 1) * NB), **mk)
    info = ds.info()
    assert info.fused == 1, "fused sm_100a kernel not selected"
    ctx = H.new_ctx(ds)
    x = x_true * 1.02
    d_x = torch.tensor(x, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(1 + NT, dtype=torch.float64, device="cuda")
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches = H.timed_device_loop(ctx, d_x, d_out, args.steps, args.warmup, sampler)
    value = world * args.steps / (ms_total * 1e-3)

    # ---- e2e: the reference-facing call, host buffers in/out, H2D + D2H inside the timed region
    G = np.empty(NT)
    nl = C.c_double()
    xh = np.ascontiguousarray(x)
    dp = H.dp
    # (argument objects built once, as any caller with a hot loop would: three ctypes conversions per call cost ~2 us)
    eval_fg, h_ctx, p_x, p_nl, p_G = L.lib.sfh_eval_fg, ctx.handle, xh.ctypes.data_as(dp), C.byref(nl), G.ctypes.data_as(dp)
    for _ in range(args.warmup):
        L.check(eval_fg(h_ctx, p_x, p_nl, p_G, None))
    H.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(H.stream)
    for _ in range(args.steps):
        L.check(eval_fg(h_ctx, p_x, p_nl, p_G, None))
    e1.record(H.stream)
    wall = time.perf_counter() - t0      # the caller-visible time of K synchronous calls (>= the event time)
    e1.synchronize()                     # (a call returns when its result packets have arrived, a little before the stream drains)
    H.barrier()
    e2e_ms = H.max_over_ranks(max(e0.elapsed_time(e1), 1e3 * wall))
    e2e_value = world * args.steps / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # sanity of what was timed: the device-resident and the host-API paths agree bit for bit
    torch.cuda.synchronize()
    out_host = d_out.cpu().numpy()
    assert np.isfinite(nl.value) and -out_host[0] == nl.value
    assert np.array_equal(out_host[1:], G)

    # ---- roofline of the dominant (fused) kernel: CUDA events around that kernel only, on its launch stream
    ms_eval, ms_kernel = C.c_double(), C.c_double()
    L.check(L.lib.sfh_time_fg(ctx.handle, xh.ctypes.data_as(dp), min(args.steps, 50), 1, 0, C.byref(ms_eval), C.byref(ms_kernel)))
    bytes_alg = NB * NT * 8 + NB * 8 + 2 * NT * 8 + 8
    achieved = bytes_alg / (ms_kernel.value * 1e-3) / 1e9
    kernel_ms_per_rank = H.gather(ms_kernel.value)     # GPUs of one box differ by a few per cent: an all-reduced step runs at the slowest
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "traffic_source": "constant: dram__bytes_read+write of this kernel from the ncu --set full capture (cold caches) summarised in profiles/ncu_summary.json (not re-measured in this run)",
                "traffic_steady_state": ncu_traffic("fused_kernel_dram_read_bytes_per_launch_steady_state"),
                "traffic_steady_state_source": "constant: dram__bytes_read of launches 21-23 of a back-to-back series, ncu application replay with caches untouched (profiles/r2_final2_warm.csv): the L2-resident head of the stack is not re-read",
                "kernel": kernel_name(info, np.float64), "kernel_ms": ms_kernel.value, "kernel_ms_per_rank": kernel_ms_per_rank,
                "bytes_alg_per_launch": bytes_alg, "peak_source": peak_src}

    # ---- the hierarchical evaluation fit_sfh / sample_sfh actually call
    fg_hier = run_hier(H, ctx, min(args.steps, 500), args.warmup, peak, bytes_alg)

    # ---- parity carried by the bench line itself
    parity = None
    no_exchange = os.environ.get("SFH_BENCH_NO_EXCHANGE") == "1"     # diagnostic: N independent shards, nothing reduced
    if world > 1 and not no_exchange:
        parity = H.parity_sharded_vs_whole(
            lambda: S.DeviceStack.synthetic(nb_total, NT, np.float64, seed=SEED, scale=1e-5, x_true=x_true, device=local), x, d_out, NT)
        assert parity["ok"], parity
    sharding = ("bin rows, one 60000-bin shard per GPU; [logL,G] (2401 f64) all-reduced every step: " + H.comm_mode(ctx)) if world > 1 else "single GPU"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        Mh, dh = ds.download()
        Mh = np.asfortranarray(Mh)
        best, routes = time_cpu(Mh, dh, x)
        t_med, n_cpu, cores = routes[best]
        other = [k for k in routes if k != best][0]
        cpu = {"value": 1.0 / t_med, "unit": UNIT, "cores": cores, "kind": "port", "route": best,
               "blas_one_thread": time_cpu_one_thread(Mh, dh, x),
               "sample": f"{n_cpu} full evaluations of the same 60000x2400 F64 stack (downloaded from the GPU), median; port of the "
                         f"reference's two-pass fg! (julia not installed): {CPU_ROUTE_TEXT[best]}; the slower restatement ({other}: "
                         f"{CPU_ROUTE_TEXT[other]}) reached {1.0 / routes[other][0]:.1f} evals/s"}
        # the CPU leg doubles as the checker of what was timed: ALL 2400 gradient components and logL against the oracle
        nlo, Go, _ = O.fg(x, Mh, dh)
        gscale = np.abs(Mh).T @ np.abs(1.0 - dh / np.maximum(Mh @ x, np.finfo(np.float64).eps))
        parity = {"vs_oracle_logl_rel": float(abs(nl.value - nlo) / abs(nlo)),
                  "vs_oracle_grad_max_err_over_scale": float(np.max(np.abs(G - Go) / gscale)), "components": NT}
        assert parity["vs_oracle_logl_rel"] <= 1e-12 and parity["vs_oracle_grad_max_err_over_scale"] <= 1e-10, parity
        del Mh
    ctx.close(); ds.close()
    del ctx, ds
    torch.cuda.empty_cache()

    config5 = None if (args.no_config5 or no_exchange) else run_config5(H, args, peak)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "per_gpu_stack_bytes": int(info.stack_bytes),
                           "l2": "inputs (1.15 GB/GPU) larger than L2 (126 MB): no flush between iterations; the stack is streamed with the L2 "
                                 f"evict_first policy except the head of every CTA's tile sequence ({int(info.l2_resident_mb)} MB in total = "
                                 f"{100.0 * info.l2_resident_mb * 1048576 / info.stack_bytes:.1f} % of the stack, loaded evict_last), which "
                                 "therefore stays L2-resident between evaluations of the same stack (SFH_L2_KEEP_MB=0 switches it off)",
                           "l2_resident_mb": int(info.l2_resident_mb),
                           "sharding": sharding,
                           "timing": "CUDA events on the launch stream; ranks aligned on the device by two untimed all-reduced steps before e0; max over ranks",
                           "value_unit_note": "N>1: value = N shard-evaluations per all-reduced step / time (weak scaling)", **tiling(info)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": NT * 8 + 8, "d2h_bytes_per_step": (NT + 1) * 16,
                        "ms_per_step": e2e_ms / args.steps,
                        "transport": "sfh_eval_fg(host coeffs) -> host [-logL, G]: an upload kernel pulls the coefficients (+ an 8-byte epoch) "
                                     "from the pinned input buffer, the finalize kernel stores every result as a 16-byte self-validating "
                                     "packet into pinned host memory, the call returns when the host has read all of them"},
                "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "fg_hier": fg_hier}
        if parity is not None:
            line["parity"] = parity
        if config5 is not None:
            line["config5"] = config5
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
