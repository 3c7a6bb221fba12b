"""End-to-end runs of the five BASELINE.json configurations through the host drivers (sfh_b200.solvers), with the
same driver running on the CPU oracle beside it where that finishes in bounded time.  Prints one JSON object per config."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
import oracle as O
from scipy import optimize

def out(**kw): print(json.dumps(kw), flush=True)

# ---- config 1: fit_templates_lbfgsb, 100x100 bins, 100 random nonnegative F64 templates, Poisson data ---------------
rng = np.random.Generator(np.random.Philox(58392))
M = np.asfortranarray(rng.random((10000, 100))); x = 100 * rng.random(100); data = rng.poisson(M @ x).astype(np.float64)
ds = S.DeviceStack(M, data)
S.fit_templates_lbfgsb(ds, data, x0=np.ones(100))       # warm-up (graph capture etc.)
t0 = time.perf_counter(); f_gpu, x_gpu = S.fit_templates_lbfgsb(ds, data, x0=np.ones(100)); t_gpu = time.perf_counter() - t0
nfev = [0]
def fg_cpu(z):
    nfev[0] += 1
    f, G, _ = O.fg(z, M, data); return float(f), G
x0r = np.ones(100) * data.sum() / (M @ np.ones(100)).sum()
t0 = time.perf_counter(); x_cpu, f_cpu, _ = optimize.fmin_l_bfgs_b(fg_cpu, x0r, bounds=[(0, None)] * 100, factr=1e-12, pgtol=1e-5, m=10, maxfun=100000, maxiter=100000); t_cpu = time.perf_counter() - t0
out(config=1, what="fit_templates_lbfgsb 100x100 bins x 100 templates F64", gpu_s=t_gpu, cpu_oracle_1thread_s=t_cpu, fevals=nfev[0],
    rel_diff_coeffs=float(np.linalg.norm(x_gpu - x_cpu) / np.linalg.norm(x_cpu)), rel_diff_nlogL=float(abs(f_gpu - f_cpu) / abs(f_cpu)),
    rel_err_vs_truth=float(np.linalg.norm(x_gpu - x) / np.linalg.norm(x)))

# ---- config 2: mcmc_sample ensemble, 1024 walkers, 200x200 bins x 500 templates F64 -----------------------------------
rng = np.random.Generator(np.random.Philox(58393))
x2 = 100 * rng.random(500)
ds2 = S.DeviceStack.synthetic(40000, 500, np.float64, 58393, 1.0, x2)
X0 = np.maximum(0.0, x2[:, None] + rng.standard_normal((500, 1024)))
d2 = ds2.download_data()
S.mcmc_sample(ds2, d2, X0, 2, rng=np.random.default_rng(1))
nsteps = 20
t0 = time.perf_counter(); chain, lps, acc = S.mcmc_sample(ds2, d2, X0, nsteps, rng=np.random.default_rng(1)); t_gpu = time.perf_counter() - t0
M2, d2 = ds2.download()
t0 = time.perf_counter(); O.mcmc_logl(X0[:, :64], M2, d2); t_cpu64 = time.perf_counter() - t0
# what the reference does NOT do but a CPU could: all walkers of a batch in one OpenBLAS GEMM (M @ X), then the Poisson sum per walker
def _cpu_gemm_logl(Xw):
    m = np.maximum(M2 @ Xw, np.finfo(np.float64).eps)
    dd = d2[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(dd > 0, dd - m - dd * np.log(dd / m), -m).sum(axis=0)
_cpu_gemm_logl(X0[:, :32])
t0 = time.perf_counter(); _cpu_gemm_logl(X0[:, :256]); t_cpu_gemm = time.perf_counter() - t0
out(config=2, what="mcmc_sample 1024 walkers x %d steps, 200x200 bins x 500 templates F64" % nsteps, gpu_s=t_gpu,
    gpu_walker_evals_per_s=1024 * nsteps / t_gpu, acceptance=acc,
    cpu_oracle_walker_evals_per_s=64 / t_cpu64, cpu_threads=O.num_threads(), cpu_sample="64 walkers, threaded over walkers (one pass over the stack per walker, as the reference's per-walker MCMCModel call)",
    cpu_batched_gemm_walker_evals_per_s=256 / t_cpu_gemm, cpu_batched_note="256 walkers in one OpenBLAS GEMM + numpy Poisson sum: NOT what the reference does (mcmc_sample.jl:12-23 is one gemv per walker)")
del M2

# ---- configs 3/4: fit_sfh PowerLawMZR + GaussianDispersion, 60 ages x 40 [M/H] = 2400 templates, 200x300 bins ------
rng = np.random.Generator(np.random.Philox(94823))
uA = np.linspace(10.1, 6.6, 60); uM = np.linspace(-2.5, 0.0, 40)
la = np.repeat(uA, 40); mh = np.tile(uM, 60)
R = rng.random(60) * 1e6
mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
xt = S.calculate_coeffs(mz, dp, R, la, mh)
ds3 = S.DeviceStack.synthetic(60000, 2400, np.float64, 94823, 1e-5, xt)
truth = np.concatenate([R, [1.0, -2.0, 0.2]])
d3 = ds3.download_data()
t0 = time.perf_counter()
res = S.fit_sfh(S.PowerLawMZR(1.2, -2.2, 6.0), S.GaussianDispersion(0.25), ds3, d3, la, mh, x0=R * 1.5, g_abstol=1e-6)
t_gpu = time.perf_counter() - t0
z = np.abs(res["map"].mu - truth) / res["map"].sigma
out(config=3, what="fit_sfh MZR (MAP+MLE BFGS), 200x300 bins x 2400 templates F64, Poisson data", gpu_s=t_gpu,
    fevals_map=int(res["map"].result.nfev), fevals_mle=int(res["mle"].result.nfev),
    frac_params_within_3sigma=float(np.mean(z < 3)), alpha_beta_sigma=[float(v) for v in res["mle"].mu[-3:]])
# config 4: the HMC inner loop = sequential logdensity_and_gradient calls on the same stack
opt = S.HierarchicalOptimizer(mz, dp, ds3, d3, la, mh, True, True, True)
xv = np.concatenate([np.log(R), [0.0, -2.0, np.log(0.2)]])
for _ in range(20): opt.logdensity_and_gradient(xv)
t0 = time.perf_counter()
for _ in range(500): opt.logdensity_and_gradient(xv)
t_leap = (time.perf_counter() - t0) / 500
Mh, dh = ds3.download()
t0 = time.perf_counter(); O.fg_hier(O.POWERLAW_MZR, (6.0,), (1, 1, 1), truth, Mh, dh, la, mh); t_cpu = time.perf_counter() - t0
out(config=4, what="HMC leapfrog = HierarchicalOptimizer.logdensity_and_gradient on the 2400-template MZR stack", gpu_us_per_eval=t_leap * 1e6,
    gpu_evals_per_s=1 / t_leap, cpu_oracle_1thread_s_per_eval=t_cpu)

# ---- config 4: tsample_sfh on the 2400-template MZR stack: short chains served by one sfh_eval_fg_hier_batched pass per round ----
tt = {}
for batched in (False, True):
    t0 = time.perf_counter()
    r = S.tsample_sfh(res, ds3, d3, la, mh, 160, eps=0.05, rng=np.random.default_rng(4), chain_length=20, max_depth=5, batched=batched)
    tt[batched] = time.perf_counter() - t0
out(config=4, what="tsample_sfh: 8 chains x 20 NUTS draws (max_depth 5) on the 2400-template MZR stack, dim 63", sequential_s=tt[False],
    batched_s=tt[True], speedup=tt[False] / tt[True], posterior_shape=list(r["posterior_matrix"].shape))
tt = {}
for batched in (False, True):
    t0 = time.perf_counter()
    r = S.tsample_sfh(res, ds3, d3, la, mh, 640, eps=0.05, rng=np.random.default_rng(4), chain_length=20, max_depth=5, batched=batched)
    tt[batched] = time.perf_counter() - t0
out(config=4, what="tsample_sfh: 32 chains x 20 NUTS draws (max_depth 5) on the 2400-template MZR stack, dim 63", sequential_s=tt[False],
    batched_s=tt[True], speedup=tt[False] / tt[True])

# ---- multi-chain hmc_sample: chains on threads sharing one sfh_eval_fg_batched pass vs chains one after another ------
from sfh_b200 import solvers as V
for (nbh, nth, nchs, nst, md) in ((60000, 200, (8,), 40, 5), (60000, 2400, (4, 8, 16), 12, 4)):
    xh = 100 * np.random.default_rng(5).random(nth)
    dsh = S.DeviceStack.synthetic(nbh, nth, np.float64, 5, 1.0, xh)
    dh2 = dsh.download_data()
    for nch in nchs:
        tt = {}
        for batched in (False, True):
            t0 = time.perf_counter()
            V.hmc_sample(dsh, dh2, nst, nchains=nch, nwarmup=nst, rng=np.random.default_rng(1), x0=xh, max_depth=md, batched=batched)
            tt[batched] = time.perf_counter() - t0
        out(what=f"hmc_sample (coroutine chains) {nbh} bins x {nth} templates F64, {nst} warm-up + {nst} draws per chain, max_depth {md}", nchains=nch,
            sequential_s=tt[False], batched_s=tt[True], speedup=tt[False] / tt[True])
    del dsh
