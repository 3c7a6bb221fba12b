// sfh_nuts.h -- native multi-chain No-U-Turn sampler around ONE batched log-density call per round
// ("batched multi-chain leapfrog ... one fg over C coefficient vectors").
//
// The reference runs its HMC chains on separate threads, every chain evaluating its own fg!
// (src/fitting/hmc_sample.jl:123-141: Threads.@threads + DynamicHMC.mcmc_with_warmup per chain;
//  src/fitting/hierarchical/generic_fitting.jl:617-626: one task per short chain in tsample_sfh).
// Here every chain also is a host thread running the textbook recursion (Hoffman & Gelman 2014, algorithm 6: slice NUTS
// with dual-averaging step-size adaptation, Gaussian kinetic energy with identity / diagonal / dense M^-1 -- the role
// DynamicHMC plays in the reference), but a chain that needs a log-density and gradient parks its request, and when every
// live chain is parked the calling thread serves them all with one batched evaluation (sfh_eval_fg_batched /
// sfh_eval_fg_hier_batched on the device: one pass over the template stack for all chains).  A chain's answers do not depend on
// how requests were grouped, so each follows exactly the trajectory it would follow alone.
//
// Random numbers: Philox4x32-10 (the generator of the device ensemble sampler, csrc/sfh_small.cuh) with key = seed,
// counter = (chain << 40 | draw index), stream 32; normals by Box-Muller.  The algorithm and the ORDER of the draws are those of
// `nuts_chain` in the Python host mirror (solvers.py), which is therefore its executable restatement: driven by the same
// Philox stream it reproduces the native chains (tests/test_native_nuts.py).  Sampler trajectories are not a parity claim
// against the reference (its engine is third-party and its tests pin shapes only).
//
// Host-only C++.
#ifndef SFH_NUTS_H
#define SFH_NUTS_H

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

namespace sfh {
namespace nuts {

// Theta: n x C column-major -> logp[C], grad n x C column-major.  Non-zero return aborts the run with that status.
using BatchLogDensity = std::function<int(const double *Theta, int64_t C, double *logp, double *grad)>;

struct Options {
    int max_depth = 8;
    int64_t nwarmup = 200;
    double delta = 0.8;   // target mean acceptance of the dual averaging
    double eps0 = 0.0;    // > 0: fixed initial step size (the reference's epsilon, generic_fitting.jl:479-482); else the doubling heuristic
    uint64_t seed = 0;
};

struct Stats {
    int64_t n_batches = 0, n_evals = 0;
};

// ---- Philox4x32-10, bit-identical to csrc/sfh_small.cuh and tests/ensemble_ref.py ------------------------------
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
inline double philox_u01(uint64_t idx, uint64_t seed, uint32_t stream) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t bits = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}
constexpr uint32_t kDrawNuts = 32;

struct Rng {
    uint64_t seed, base, n = 0;
    Rng(uint64_t s, uint64_t chain) : seed(s), base(chain << 40) {}
    double random() { return philox_u01(base | n++, seed, kDrawNuts); }
    double normal() {
        const double u1 = random(), u2 = random();
        return std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(6.283185307179586 * u2);
    }
};

// ---- Gaussian kinetic energy: M^-1 = identity (kind 0), diag(inv_mass) (1) or the dense matrix inv_mass (2) --------
struct Mass {
    int kind = 0;
    int64_t n = 0;
    std::vector<double> inv, Lc;   // Lc: lower Cholesky factor of the symmetrised dense inv_mass (column-major)
    bool init(int kind_, int64_t n_, const double *inv_mass) {
        kind = kind_; n = n_;
        if (kind == 1) inv.assign(inv_mass, inv_mass + n);
        if (kind == 2) {
            inv.assign(inv_mass, inv_mass + n * n);   // M^-1 r uses the matrix as given; its symmetric part is factorised
            std::vector<double> sym((size_t)n * n);
            for (int64_t j = 0; j < n; ++j)
                for (int64_t i = 0; i < n; ++i) sym[(size_t)(i + j * n)] = 0.5 * (inv_mass[i + j * n] + inv_mass[j + i * n]);
            Lc.assign((size_t)n * n, 0.0);
            for (int64_t j = 0; j < n; ++j) {
                double d = sym[(size_t)(j + j * n)];
                for (int64_t k = 0; k < j; ++k) d -= Lc[(size_t)(j + k * n)] * Lc[(size_t)(j + k * n)];
                if (!(d > 0.0)) return false;   // not positive definite
                const double ljj = std::sqrt(d);
                Lc[(size_t)(j + j * n)] = ljj;
                for (int64_t i = j + 1; i < n; ++i) {
                    double v = sym[(size_t)(i + j * n)];
                    for (int64_t k = 0; k < j; ++k) v -= Lc[(size_t)(i + k * n)] * Lc[(size_t)(j + k * n)];
                    Lc[(size_t)(i + j * n)] = v / ljj;
                }
            }
        }
        return true;
    }
    void vel(const double *r, double *out) const {   // M^-1 r
        if (kind == 0) { std::copy(r, r + n, out); return; }
        if (kind == 1) { for (int64_t i = 0; i < n; ++i) out[i] = inv[(size_t)i] * r[i]; return; }
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) { const double rj = r[j]; const double *col = &inv[(size_t)(j * n)]; for (int64_t i = 0; i < n; ++i) out[i] += col[i] * rj; }
    }
    void draw(Rng &rng, double *r) const {           // r ~ N(0, M)
        for (int64_t i = 0; i < n; ++i) r[i] = rng.normal();
        if (kind == 1) for (int64_t i = 0; i < n; ++i) r[i] /= std::sqrt(inv[(size_t)i]);
        if (kind == 2)                               // M = inv^-1 = Lc^-T Lc^-1  =>  r = Lc^-T z: back substitution on Lc^T
            for (int64_t i = n - 1; i >= 0; --i) {
                double v = r[i];
                for (int64_t k = i + 1; k < n; ++k) v -= Lc[(size_t)(k + i * n)] * r[k];
                r[i] = v / Lc[(size_t)(i + i * n)];
            }
    }
};

struct Abort {};   // thrown inside a chain thread when the run is being torn down

// ---- one chain -------------------------------------------------------------------------------------------------------
class Chain {
public:
    using Eval = std::function<void(const double *theta, double *lp, double *grad)>;   // blocks until served; may throw Abort
    Chain(int64_t n, const Mass &mass, const Options &opt, uint64_t chain_id, Eval eval)
        : n_(n), mass_(mass), opt_(opt), rng_(opt.seed, chain_id), eval_(std::move(eval)), tmp_((size_t)n) {}

    // theta0[n]; samples: n x nsteps column-major; logps[nsteps]; returns the final step size
    double run(const double *theta0, int64_t nsteps, double *samples, double *logps) {
        const int64_t d = n_;
        Vec theta(theta0, theta0 + d), grad((size_t)d), r0((size_t)d);
        double lp;
        eval_(theta.data(), &lp, grad.data());
        double eps = 0.1 / std::sqrt((double)d);
        mass_.draw(rng_, r0.data());
        State s1 = leapfrog(theta, r0, grad, eps);
        const double H0h = lp - kin(r0);
        double H1 = s1.lp - kin(s1.r);
        const double a = (std::isfinite(H1) && H1 - H0h > std::log(0.5)) ? 1.0 : -1.0;
        for (int it = 0; it < (opt_.eps0 > 0 ? 0 : 50); ++it) {
            s1 = leapfrog(theta, r0, grad, eps);
            H1 = s1.lp - kin(s1.r);
            if (!std::isfinite(H1)) H1 = -std::numeric_limits<double>::infinity();
            if (a * (H1 - H0h) <= -a * std::log(2.0)) break;
            eps *= std::pow(2.0, a);
        }
        if (opt_.eps0 > 0) eps = opt_.eps0;
        const double mu = std::log(10 * eps), gamma = 0.05, t0 = 10.0, kappa = 0.75;
        double ebar = 1.0, Hbar = 0.0;
        for (int64_t m = 1; m <= opt_.nwarmup + nsteps; ++m) {
            mass_.draw(rng_, r0.data());
            const double H0 = lp - kin(r0);
            const double logu = H0 + std::log(rng_.random());
            Vec thm = theta, thp = theta, rm = r0, rp = r0, gm = grad, gp = grad;
            int j = 0, sflag = 1;
            double nn = 1;
            double alpha = 0.0; int nalpha = 1;
            while (sflag && j < opt_.max_depth) {
                const int v = rng_.random() < 0.5 ? -1 : 1;
                Tree t = v == -1 ? build(thm, rm, gm, logu, v, j, eps, H0) : build(thp, rp, gp, logu, v, j, eps, H0);
                if (v == -1) { thm = t.thm; rm = t.rm; gm = t.gm; } else { thp = t.thp; rp = t.rp; gp = t.gp; }
                alpha = t.a; nalpha = t.na;
                if (t.s && rng_.random() < std::min(1.0, t.n / nn)) { theta = t.th1; lp = t.lp1; grad = t.g1; }
                nn += t.n;
                sflag = t.s * uturn_ok(thm, thp, rm, rp);
                ++j;
            }
            if (m <= opt_.nwarmup) {   // dual averaging
                Hbar = (1.0 - 1.0 / (m + t0)) * Hbar + (opt_.delta - alpha / nalpha) / (m + t0);
                const double leps = mu - std::sqrt((double)m) / gamma * Hbar;
                const double eta = std::pow((double)m, -kappa);
                ebar = std::exp(eta * leps + (1 - eta) * std::log(ebar));
                eps = std::exp(leps);
                if (m == opt_.nwarmup) eps = ebar;
            } else {
                const int64_t k = m - opt_.nwarmup - 1;
                std::copy(theta.begin(), theta.end(), samples + k * d);
                logps[k] = lp;
            }
        }
        return eps;
    }

private:
    using Vec = std::vector<double>;
    struct State { Vec th, r, g; double lp; };
    struct Tree { Vec thm, rm, gm, thp, rp, gp, th1, g1; double lp1 = 0, n = 0, a = 0; int s = 0, na = 0; };

    double dot(const Vec &a, const Vec &b) const { double s = 0; for (int64_t i = 0; i < n_; ++i) s += a[(size_t)i] * b[(size_t)i]; return s; }
    double kin(const Vec &r) { mass_.vel(r.data(), tmp_.data()); return 0.5 * dot(r, tmp_); }
    int uturn_ok(const Vec &thm, const Vec &thp, const Vec &rm, const Vec &rp) {
        Vec dth((size_t)n_);
        for (int64_t i = 0; i < n_; ++i) dth[(size_t)i] = thp[(size_t)i] - thm[(size_t)i];
        mass_.vel(rm.data(), tmp_.data());
        const int a = dot(dth, tmp_) >= 0;
        mass_.vel(rp.data(), tmp_.data());
        const int b = dot(dth, tmp_) >= 0;
        return a * b;
    }
    State leapfrog(const Vec &theta, const Vec &r, const Vec &grad, double eps) {
        State s;
        s.r.resize((size_t)n_); s.th.resize((size_t)n_); s.g.resize((size_t)n_);
        for (int64_t i = 0; i < n_; ++i) s.r[(size_t)i] = r[(size_t)i] + 0.5 * eps * grad[(size_t)i];
        mass_.vel(s.r.data(), tmp_.data());
        for (int64_t i = 0; i < n_; ++i) s.th[(size_t)i] = theta[(size_t)i] + eps * tmp_[(size_t)i];
        eval_(s.th.data(), &s.lp, s.g.data());
        for (int64_t i = 0; i < n_; ++i) s.r[(size_t)i] += 0.5 * eps * s.g[(size_t)i];
        return s;
    }
    Tree build(const Vec &theta, const Vec &r, const Vec &grad, double logu, int v, int j, double eps, double H0) {
        if (j == 0) {
            State st = leapfrog(theta, r, grad, v * eps);
            double Hn = st.lp - kin(st.r);
            if (!std::isfinite(Hn)) Hn = -std::numeric_limits<double>::infinity();
            Tree t;
            t.n = logu <= Hn ? 1 : 0;
            t.s = logu < Hn + 1000.0 ? 1 : 0;
            t.a = std::min(1.0, std::exp(std::min(0.0, Hn - H0)));
            t.na = 1;
            t.lp1 = st.lp;
            t.thm = st.th; t.thp = st.th; t.th1 = st.th; t.rm = st.r; t.rp = st.r; t.gm = st.g; t.gp = st.g; t.g1 = std::move(st.g);
            return t;
        }
        Tree t = build(theta, r, grad, logu, v, j - 1, eps, H0);
        if (t.s) {
            Tree t2 = v == -1 ? build(t.thm, t.rm, t.gm, logu, v, j - 1, eps, H0) : build(t.thp, t.rp, t.gp, logu, v, j - 1, eps, H0);
            if (v == -1) { t.thm = std::move(t2.thm); t.rm = std::move(t2.rm); t.gm = std::move(t2.gm); }
            else { t.thp = std::move(t2.thp); t.rp = std::move(t2.rp); t.gp = std::move(t2.gp); }
            if (t.n + t2.n > 0 && rng_.random() < t2.n / (t.n + t2.n)) { t.th1 = std::move(t2.th1); t.lp1 = t2.lp1; t.g1 = std::move(t2.g1); }
            t.s = t2.s * uturn_ok(t.thm, t.thp, t.rm, t.rp);
            t.n += t2.n; t.a += t2.a; t.na += t2.na;
        }
        return t;
    }

    int64_t n_;
    const Mass &mass_;
    Options opt_;
    Rng rng_;
    Eval eval_;
    Vec tmp_;
};

// ---- many chains, one batched evaluation per round -----------------------------------------------------------------------
// theta0: n x nchains column-major; nsteps[c] draws for chain c; samples: n x sum(nsteps) column-major, chains concatenated in
// order; logps[sum(nsteps)]; step_sizes[nchains].  Returns 0, or the batch function's status.
inline int run_chains(const BatchLogDensity &fn, int64_t n, int64_t nchains, const double *theta0, const int64_t *nsteps,
                      const Mass &mass, const Options &opt, double *samples, double *logps, double *step_sizes, Stats *stats) {
    struct Slot { const double *theta = nullptr; double *lp = nullptr, *grad = nullptr; bool waiting = false, served = false; };
    std::vector<Slot> slots((size_t)nchains);
    std::mutex mu;
    std::condition_variable cv_req, cv_done;
    int64_t live = nchains, pending = 0;
    bool abort = false;
    int status = 0;
    std::vector<int64_t> offs((size_t)nchains + 1, 0);
    for (int64_t c = 0; c < nchains; ++c) offs[(size_t)c + 1] = offs[(size_t)c] + nsteps[c];

    std::vector<std::thread> threads;
    threads.reserve((size_t)nchains);
    for (int64_t c = 0; c < nchains; ++c) {
        threads.emplace_back([&, c] {
            auto eval = [&, c](const double *theta, double *lp, double *grad) {
                std::unique_lock<std::mutex> lk(mu);
                if (abort) throw Abort{};
                Slot &s = slots[(size_t)c];
                s.theta = theta; s.lp = lp; s.grad = grad; s.waiting = true; s.served = false;
                ++pending;
                cv_req.notify_one();
                cv_done.wait(lk, [&] { return s.served || abort; });
                if (!s.served) throw Abort{};
            };
            try {
                Chain ch(n, mass, opt, (uint64_t)c, eval);
                step_sizes[c] = ch.run(theta0 + c * n, nsteps[c], samples + offs[(size_t)c] * n, logps + offs[(size_t)c]);
            } catch (const Abort &) {
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu);
                if (!status) status = -1;
                abort = true;
            }
            std::lock_guard<std::mutex> lk(mu);
            --live;
            cv_req.notify_one();
            cv_done.notify_all();
        });
    }
    std::vector<double> Theta, lp, grad;
    std::vector<int64_t> ids;
    {
        std::unique_lock<std::mutex> lk(mu);
        while (true) {
            cv_req.wait(lk, [&] { return pending == live || abort; });
            if (live == 0 || abort) break;
            ids.clear();
            for (int64_t c = 0; c < nchains; ++c) if (slots[(size_t)c].waiting) ids.push_back(c);
            const int64_t C = (int64_t)ids.size();
            Theta.resize((size_t)(n * C)); lp.resize((size_t)C); grad.resize((size_t)(n * C));
            for (int64_t k = 0; k < C; ++k) std::copy(slots[(size_t)ids[(size_t)k]].theta, slots[(size_t)ids[(size_t)k]].theta + n, Theta.data() + k * n);
            lk.unlock();
            const int st = fn(Theta.data(), C, lp.data(), grad.data());
            lk.lock();
            if (st != 0) { status = st; abort = true; cv_done.notify_all(); break; }
            if (stats) { ++stats->n_batches; stats->n_evals += C; }
            for (int64_t k = 0; k < C; ++k) {
                Slot &s = slots[(size_t)ids[(size_t)k]];
                *s.lp = lp[(size_t)k];
                std::copy(grad.data() + k * n, grad.data() + (k + 1) * n, s.grad);
                s.waiting = false; s.served = true;
            }
            pending = 0;
            cv_done.notify_all();
        }
        if (abort) cv_done.notify_all();
    }
    for (auto &t : threads) t.join();
    return status;
}

// The log-density sample_sfh / tsample_sfh hand to DynamicHMC: LogDensityProblems.logdensity_and_gradient(::HierarchicalOptimizer,
// xvec) (fitting/hierarchical/generic_fitting.jl:90-199; jacobian corrections on, :477) for C chains at once, around an `inner`
// batched hierarchical fg! over the natural variables (V: (nj+npar) x C -> -logL[C], gradient (nj+npar) x C).
// X: (nj + nfree) x C transformed free variables -> (+logp[C], +gradient).
using InnerBatched = std::function<int(const double *V, int64_t C, double *neg_logl, double *G)>;
inline BatchLogDensity hier_logdensity_batched(InnerBatched inner, int64_t nj, int npar, const double *params0, const int32_t *transforms,
                                               const uint8_t *free_mask, bool jacobian_corrections) {
    std::vector<double> p0(params0, params0 + npar);
    std::vector<int32_t> tf(transforms, transforms + npar);
    std::vector<uint8_t> fr(free_mask, free_mask + npar);
    int nfree = 0;
    for (int k = 0; k < npar; ++k) nfree += fr[(size_t)k] ? 1 : 0;
    const int64_t nv = nj + npar, nx = nj + nfree;
    return [=](const double *X, int64_t C, double *logp, double *grad) -> int {
        std::vector<double> V((size_t)(nv * C)), G((size_t)(nv * C));
        for (int64_t c = 0; c < C; ++c) {
            const double *xv = X + c * nx;
            double *v = V.data() + c * nv;
            for (int64_t i = 0; i < nj; ++i) v[i] = std::exp(xv[i]);                                       // :127
            for (int k = 0, q = 0; k < npar; ++k) {
                if (!fr[(size_t)k]) { v[nj + k] = p0[(size_t)k]; continue; }                               // :134-136
                const double t = xv[nj + q++];
                v[nj + k] = tf[(size_t)k] == 1 ? std::exp(t) : t;                                          // :129-131
            }
        }
        const int st = inner(V.data(), C, logp, G.data());                                                // :140 (logp holds -logL here)
        if (st) return st;
        for (int64_t c = 0; c < C; ++c) {
            const double *xv = X + c * nx, *v = V.data() + c * nv, *g2 = G.data() + c * nv;
            double *g = grad + c * nx;
            double nl = logp[c];
            for (int64_t i = 0; i < nj; ++i) {
                if (jacobian_corrections) { nl -= xv[i]; g[i] = -(g2[i] * v[i] - 1.0); }                    // :150-154
                else g[i] = -(g2[i] * v[i]);                                                              // :163-164
            }
            for (int k = 0, q = 0; k < npar; ++k) {
                if (!fr[(size_t)k]) continue;                                                            // :181-189
                double gk = g2[nj + k];
                if (tf[(size_t)k] == 1) {
                    gk *= v[nj + k];
                    if (jacobian_corrections) { nl -= std::log(v[nj + k]); gk -= 1.0; }
                }
                g[nj + q++] = -gk;
            }
            logp[c] = -nl;                                                                               // :193-197
        }
        return 0;
    };
}

}  // namespace nuts
}  // namespace sfh
#endif  // SFH_NUTS_H
