// sfh_drivers.h -- native host loops that OWN the iteration around the device evaluations (SURVEY.md section 8f rank 1,
// DESIGN.md section 7 item 2): a dense BFGS with a strong-Wolfe line search, the engine behind fit_templates
// (src/fitting/solvers.jl:163-221), fit_templates_fast (:238-275), fit_sfh (fitting/hierarchical/generic_fitting.jl:242-411)
// and fixed_amr (hierarchical/fixed_amr.jl:166-167), all of which call
//       Optim.optimize(only_fg!(...), x0, BFGS(alphaguess = InitialStatic(1.0, true), linesearch = HagerZhang()), ...)
// and read the final inverse-Hessian estimate off the trace for the parameter uncertainties.
//
// The optimisation ENGINE is third-party in the reference (Optim.jl / LineSearches.jl, un-vendored): this is a textbook
// restatement of the same method (Nocedal & Wright, Numerical Optimization 2nd ed.: BFGS update eq. 6.17, line search
// algorithms 3.5 / 3.6 with safeguarded cubic interpolation), with the reference's InitialStatic(1.0, scaled = true) first
// trial step, the inf-norm gradient stopping rule of Optim's `g_abstol`, and the inverse Hessian returned to the caller.
// Iterates differ between engines; converged answers do not (tests/test_native_bfgs.py).
//
// Host-only C++; the objective is a callback, so the same loop serves a device-bound fg! (sfh_api.cu) or any C function.
#ifndef SFH_DRIVERS_H
#define SFH_DRIVERS_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace sfh {
namespace drivers {

// f(x) and its gradient in one call (Optim's only_fg! protocol); non-zero return aborts the run with that status
using Objective = std::function<int(const double *x, double *f, double *g)>;

struct BfgsOptions {
    double g_abstol = 1e-8;   // stop when max_i |g_i| <= g_abstol (Optim.Options g_abstol; solvers.jl:206, generic_fitting.jl:304)
    int64_t maxiter = 5000;   // solvers.jl:203 / generic_fitting.jl: iterations kwarg
    double c1 = 1e-4, c2 = 0.9;
    int alphaguess = 1;       // 1: InitialStatic(1.0, scaled = true): first trial step of length min(1, |p|) (the reference);
                              // 0: from the previous decrease, 2 (f_k - f_{k-1}) / phi'(0), capped at 1 (Nocedal & Wright p. 59)
    int max_linesearch = 40;  // objective evaluations per line search before it is declared failed
};

struct BfgsReport {
    double f = 0.0, g_norm = 0.0;
    int64_t iterations = 0, f_calls = 0;
    int converged = 0;        // 1: g_norm <= g_abstol
    int status = 0;           // 0 converged; 1 iteration limit; 2 line search failed (precision loss); 3 objective not finite at x0
};

namespace detail {
inline double dot(const double *a, const double *b, int64_t n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) { s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3]; }
    for (; i < n; ++i) s0 += a[i] * b[i];
    return (s0 + s1) + (s2 + s3);
}
inline double infnorm(const double *a, int64_t n) {
    double m = 0;
    for (int64_t i = 0; i < n; ++i) { const double v = std::fabs(a[i]); if (!(v <= m)) m = v; }   // NaN propagates
    return m;
}

// columns [j0, j1) of an n x n column-major matrix, split over host threads when the matrix is large (T = 2400: 46 MB)
template <typename F>
inline void for_columns(int64_t n, F &&body) {
    const int64_t work = n * n;
    int nth = 1;
    if (work >= (int64_t)1 << 18) nth = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), std::min<int64_t>(16, n / 64));
    if (nth <= 1) { body((int64_t)0, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + nth - 1) / nth;
    for (int t = 1; t < nth; ++t) {
        const int64_t j0 = per * t, j1 = std::min(n, j0 + per);
        if (j0 < j1) th.emplace_back([=, &body] { body(j0, j1); });
    }
    body((int64_t)0, std::min(n, per));
    for (auto &x : th) x.join();
}

// minimiser of the cubic through (a, fa) with slope fpa, (b, fb), (c, fc); NaN when it does not exist
inline double cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc) {
    const double C = fpa, db = b - a, dc = c - a;
    const double denom = (db * dc) * (db * dc) * (db - dc);
    if (denom == 0.0 || !std::isfinite(denom)) return std::numeric_limits<double>::quiet_NaN();
    const double t0 = fb - fa - C * db, t1 = fc - fa - C * dc;
    const double A = (dc * dc * t0 - db * db * t1) / denom;
    const double B = (-dc * dc * dc * t0 + db * db * db * t1) / denom;
    const double rad = B * B - 3 * A * C;
    if (!(rad >= 0.0) || A == 0.0) return std::numeric_limits<double>::quiet_NaN();
    return a + (-B + std::sqrt(rad)) / (3 * A);
}
// minimiser of the parabola through (a, fa) with slope fpa and (b, fb)
inline double quadmin(double a, double fa, double fpa, double b, double fb) {
    const double db = b - a, B = (fb - fa - fpa * db) / (db * db);
    if (!(B > 0.0) || !std::isfinite(B)) return std::numeric_limits<double>::quiet_NaN();
    return a - fpa / (2.0 * B);
}
}  // namespace detail

// Strong-Wolfe line search along p from x (Nocedal & Wright algorithms 3.5 and 3.6).  On success x_new / g_new / *f_new hold
// the accepted point.  Returns 0 on success, -1 if no acceptable step was found, or the objective's own (positive) status.
inline int wolfe_search(const Objective &fn, int64_t n, const double *x, const double *p, double f0, double dphi0, double alpha1,
                        const BfgsOptions &o, double *x_new, double *g_new, double *f_new, double *alpha_out, int64_t *f_calls,
                        double amax = std::numeric_limits<double>::infinity()) {
    using namespace detail;
    int evals = 0;
    alpha1 = std::min(alpha1, amax);
    int status = 0;
    auto phi = [&](double a, double *dphi) -> double {
        for (int64_t i = 0; i < n; ++i) x_new[i] = x[i] + a * p[i];
        double f = 0;
        status = fn(x_new, &f, g_new);
        ++evals; ++*f_calls;
        *dphi = dot(g_new, p, n);
        if (!std::isfinite(f)) { f = std::numeric_limits<double>::infinity(); *dphi = std::numeric_limits<double>::quiet_NaN(); }
        return f;
    };
    // Hager & Zhang's APPROXIMATE Wolfe conditions (the reference's line search is LineSearches.HagerZhang): once the
    // objective changes by less than its own rounding noise the sufficient-decrease test is meaningless, so a trial whose
    // value is within a relative 1e-10 of f0 is accepted on its slope alone, (2 delta - 1) phi'(0) >= phi'(a) >= c2 phi'(0)
    // with delta = 0.1.  Without this the Poisson objectives stall at |g| ~ 1e-5 (f ~ 1e3 summed over 1e4+ bins).
    auto approx_wolfe = [&](double fa, double da) {
        return fa <= f0 + 1e-10 * std::fabs(f0) && da >= o.c2 * dphi0 && da <= -0.8 * dphi0;
    };
    auto zoom = [&](double lo, double hi, double flo, double fhi, double dlo, double *alpha) -> int {
        double rec = 0, frec = f0;   // the previous trial, third point of the cubic
        for (int it = 0; evals < o.max_linesearch; ++it) {
            const double d = hi - lo, a = std::min(lo, hi), b = std::max(lo, hi);
            double aj = std::numeric_limits<double>::quiet_NaN();
            if (it > 0) aj = cubicmin(lo, flo, dlo, hi, fhi, rec, frec);
            if (!(aj > a + 0.2 * std::fabs(d) && aj < b - 0.2 * std::fabs(d))) {        // too close to an end, or no minimiser
                aj = quadmin(lo, flo, dlo, hi, fhi);
                if (!(aj > a + 0.1 * std::fabs(d) && aj < b - 0.1 * std::fabs(d))) aj = lo + 0.5 * d;
            }
            double dj;
            const double fj = phi(aj, &dj);
            if (status) return status;
            if (approx_wolfe(fj, dj)) { *alpha = aj; *f_new = fj; return 0; }
            if (fj > f0 + o.c1 * aj * dphi0 || fj >= flo) {
                rec = hi; frec = fhi; hi = aj; fhi = fj;
            } else {
                if (std::fabs(dj) <= -o.c2 * dphi0) { *alpha = aj; *f_new = fj; return 0; }
                if (dj * (hi - lo) >= 0) { rec = hi; frec = fhi; hi = lo; fhi = flo; } else { rec = lo; frec = flo; }
                lo = aj; flo = fj; dlo = dj;
            }
            if (std::fabs(hi - lo) <= 1e-16 * std::max(1.0, std::fabs(lo))) break;     // interval collapsed
        }
        return -1;
    };
    // x_new / g_new always hold the LAST evaluated trial; zoom and the acceptance below return right after evaluating the
    // point they accept, so on success that is the accepted one
    double a0 = 0, fa0 = f0, da0 = dphi0, a1 = alpha1;
    for (int i = 0; evals < o.max_linesearch; ++i) {
        double da1;
        const double fa1 = phi(a1, &da1);
        if (status) return status;
        if (approx_wolfe(fa1, da1)) { *alpha_out = a1; *f_new = fa1; return 0; }
        if (fa1 > f0 + o.c1 * a1 * dphi0 || (i > 0 && fa1 >= fa0)) return zoom(a0, a1, fa0, fa1, da0, alpha_out);
        if (std::fabs(da1) <= -o.c2 * dphi0) { *alpha_out = a1; *f_new = fa1; return 0; }
        if (da1 >= 0) return zoom(a1, a0, fa1, fa0, da1, alpha_out);
        if (a1 >= amax) { *alpha_out = a1; *f_new = fa1; return 0; }   // still descending at the largest step the bounds allow: take it
        // still descending steeply: extrapolate to where the secant through the two slopes vanishes, kept inside
        // [1.25, 10] x the current step (the scaled first trial can be orders of magnitude short of the minimiser)
        double next = 10.0 * a1;
        if (da1 > da0) next = std::min(10.0 * a1, std::max(1.25 * a1, a1 - da1 * (a1 - a0) / (da1 - da0)));
        a0 = a1; fa0 = fa1; da0 = da1;
        a1 = std::min(next, amax);
    }
    return -1;
}

// Where the n x n inverse Hessian lives when it is not the host array: the three operations the loop needs from it.  The
// device implementation (sfh_api.cu: DeviceHessian) keeps a 2400-template inverse Hessian (46 MB) in HBM, where the matrix-
// vector product and the rank-two update cost ~10 us each instead of ~8 ms on the host cores.  Non-zero returns abort the run.
struct HessianBackend {
    virtual int reset_identity() = 0;
    virtual int matvec(const double *g, double *q) = 0;                                  // q = H g
    virtual int rank2(const double *s, const double *Hy, double rho, double cs) = 0;     // H += (cs s - rho Hy) s' - rho s Hy'
    virtual int download(double *invH) = 0;                                              // column-major n x n
    virtual ~HessianBackend() {}
};

// Dense BFGS on the inverse Hessian (column-major n x n in invH, which must hold n*n doubles; it is initialised to the
// identity here, as Optim does).  x: in = start, out = minimiser.  With a backend the matrix lives there during the run and
// invH (nullable then) receives its final value.
inline int bfgs_minimize(const Objective &fn, int64_t n, double *x, const BfgsOptions &o, BfgsReport *rep, double *invH,
                         HessianBackend *hb = nullptr) {
    using namespace detail;
    std::vector<double> g((size_t)n), gn((size_t)n), xn((size_t)n), p((size_t)n), s((size_t)n), y((size_t)n), Hy((size_t)n);
    // Host matrix: the rank-two update of one iteration is DEFERRED and applied column by column inside the next iteration's
    // q = H g pass (the column is updated, then dotted while it is still in cache): one read + write of the n x n matrix per
    // iteration instead of a read (product) plus a read + write (update).  Same expressions per column, same dot: identical bits.
    bool pending = false;
    double p_rho = 0, p_cs = 0;
    std::vector<double> ps, pHy;
    auto apply_pending = [&](int64_t j0, int64_t j1) {
        for (int64_t j = j0; j < j1; ++j) {
            double *col = invH + j * n;
            const double a = p_cs * ps[(size_t)j] - p_rho * pHy[(size_t)j], b = -p_rho * ps[(size_t)j];
            for (int64_t i = 0; i < n; ++i) col[i] += a * ps[(size_t)i] + b * pHy[(size_t)i];
        }
    };
    auto identity = [&]() -> int {
        pending = false;   // the matrix is being replaced: whatever was still to be added to it is void
        if (hb) return hb->reset_identity();
        for (int64_t j = 0; j < n; ++j) {
            double *col = invH + j * n;
            std::fill(col, col + n, 0.0);
            col[j] = 1.0;
        }
        return 0;
    };
    auto finish = [&](const BfgsReport &rr) -> int {
        *rep = rr;
        if (pending) { for_columns(n, apply_pending); pending = false; }
        return (hb && invH) ? hb->download(invH) : 0;
    };
    if (int e = identity()) return e;
    BfgsReport r;
    double f = 0;
    int st = fn(x, &f, g.data());
    r.f_calls = 1;
    if (st) return st;
    r.f = f; r.g_norm = infnorm(g.data(), n);
    if (!std::isfinite(f) || !std::isfinite(r.g_norm)) { r.status = 3; return finish(r); }
    double f_prev = f + std::sqrt(dot(g.data(), g.data(), n)) / 2.0;
    // One product with the n x n matrix per iteration instead of two: q = H g_new (one read) gives both H y = q + p_old
    // (p_old = -H g_old) and, after the rank-two update (one read + write), the next direction algebraically:
    //   -p_new = H_new g_new = q - rho (s (Hy.g) + Hy (s.g)) + (rho^2 y'Hy + rho) s (s.g)
    for (int64_t j = 0; j < n; ++j) p[(size_t)j] = -g[(size_t)j];   // H = I
    std::vector<double> &q = Hy;                                    // q is overwritten by Hy in place
    while (true) {
        if (r.g_norm <= o.g_abstol) { r.converged = 1; r.status = 0; break; }
        if (r.iterations >= o.maxiter) { r.status = 1; break; }
        double dphi0 = dot(g.data(), p.data(), n);
        if (!(dphi0 < 0)) {   // not a descent direction (H lost positive definiteness to rounding): restart from the identity
            if (int e = identity()) return e;
            for (int64_t j = 0; j < n; ++j) p[(size_t)j] = -g[(size_t)j];
            dphi0 = -dot(g.data(), g.data(), n);
            if (!(dphi0 < 0)) { r.status = 2; break; }
        }
        double alpha1 = 1.0;
        if (o.alphaguess == 1) {
            const double pn = std::sqrt(dot(p.data(), p.data(), n));
            alpha1 = pn > 0 ? std::min(1.0, pn) / pn : 1.0;
        } else {
            const double a = 1.01 * 2.0 * (f - f_prev) / dphi0;
            alpha1 = (a > 0 && std::isfinite(a)) ? std::min(1.0, a) : 1.0;
        }
        double alpha = 0, fnew = f;
        st = wolfe_search(fn, n, x, p.data(), f, dphi0, alpha1, o, xn.data(), gn.data(), &fnew, &alpha, &r.f_calls);
        if (st == -1) { r.status = 2; break; }
        if (st) return st;
        for (int64_t i = 0; i < n; ++i) { s[(size_t)i] = xn[(size_t)i] - x[i]; y[(size_t)i] = gn[(size_t)i] - g[(size_t)i]; x[i] = xn[(size_t)i]; }
        g.swap(gn);
        f_prev = f; f = fnew;
        ++r.iterations;
        r.f = f; r.g_norm = infnorm(g.data(), n);
        if (r.g_norm <= o.g_abstol) { r.converged = 1; r.status = 0; break; }
        if (!std::isfinite(f)) { r.status = 2; break; }
        // H <- (I - rho s y') H (I - rho y s') + rho s s'  =  H - rho (s Hy' + Hy s') + (rho^2 y'Hy + rho) s s'
        if (hb) {
            if (int e = hb->matvec(g.data(), q.data())) return e;
        } else {
            for_columns(n, [&](int64_t j0, int64_t j1) {
                for (int64_t j = j0; j < j1; ++j) {
                    if (pending) apply_pending(j, j + 1);
                    q[(size_t)j] = dot(invH + j * n, g.data(), n);   // H symmetric: column j = row j
                }
            });
            pending = false;
        }
        const double ys = dot(y.data(), s.data(), n);
        if (!(ys > 0)) {      // curvature condition violated (cannot happen with a Wolfe step up to rounding): skip the update
            for (int64_t j = 0; j < n; ++j) p[(size_t)j] = -q[(size_t)j];
            continue;
        }
        const double rho = 1.0 / ys;
        for (int64_t j = 0; j < n; ++j) { const double qj = q[(size_t)j]; Hy[(size_t)j] = qj + p[(size_t)j]; p[(size_t)j] = -qj; }   // p holds -q for now
        const double yHy = dot(y.data(), Hy.data(), n);
        const double cs = rho * rho * yHy + rho;
        const double Hyg = dot(Hy.data(), g.data(), n), sg = dot(s.data(), g.data(), n);
        if (hb) {
            if (int e = hb->rank2(s.data(), Hy.data(), rho, cs)) return e;
        } else {
            ps = s; pHy = Hy; p_rho = rho; p_cs = cs;   // applied by the next q = H g pass, or by finish()
            pending = true;
        }
        for (int64_t j = 0; j < n; ++j) p[(size_t)j] += rho * (s[(size_t)j] * Hyg + Hy[(size_t)j] * sg) - cs * s[(size_t)j] * sg;
    }
    return finish(r);
}


// ---------------------------------------------------------------------------------------------------------------------
// L-BFGS-B: fit_templates_lbfgsb (src/fitting/solvers.jl:70-90) hands fg! to LBFGSB.jl (the Fortran code of Zhu, Byrd, Lu &
// Nocedal) with lb = 0, ub = Inf, m = 10, factr = 1e-12, pgtol = 1e-5.  Restated from the published algorithm (Byrd, Lu,
// Nocedal & Zhu, SIAM J. Sci. Comput. 16 (1995): compact representation B = theta I - W M W', generalized Cauchy point
// (algorithm CP), direct primal subspace minimisation (section 5.1) with the projection refinement of Morales & Nocedal
// (ACM TOMS 38 (2011)), line search limited to the feasible segment, curvature-checked updates, projected-gradient and
// relative-reduction stopping rules).  Third-party engine in the reference: converged answers are pinned, iterates are not.
// ---------------------------------------------------------------------------------------------------------------------
struct LbfgsbOptions {
    int m = 10;               // stored correction pairs (solvers.jl:82)
    double factr = 1e-12;     // stop when (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= factr * eps(Float64): in units of machine epsilon
                              // like the Fortran code, so the reference's 1e-12 (solvers.jl:82) leaves only "no change at all"
    double pgtol = 1e-5;      // stop when the infinity norm of the projected gradient <= pgtol
    int64_t maxiter = 100000, maxfun = 100000;
};
struct LbfgsbReport {
    double f = 0, pg_norm = 0;
    int64_t iterations = 0, f_calls = 0;
    int status = 0;           // 0: projected gradient <= pgtol; 1: relative reduction <= factr; 2: iteration / evaluation limit;
                              // 3: line search failed twice (abnormal termination); 4: start not finite
};

namespace detail {
// solve A X = B in place for small dense systems (column-major A: k x k, B: k x nrhs), partial pivoting; false if singular
inline bool solve_small(std::vector<double> A, int k, std::vector<double> &B, int nrhs) {
    for (int c = 0; c < k; ++c) {
        int piv = c;
        for (int r = c + 1; r < k; ++r) if (std::fabs(A[(size_t)(r + c * k)]) > std::fabs(A[(size_t)(piv + c * k)])) piv = r;
        if (A[(size_t)(piv + c * k)] == 0.0) return false;
        if (piv != c) {
            for (int j = 0; j < k; ++j) std::swap(A[(size_t)(c + j * k)], A[(size_t)(piv + j * k)]);
            for (int j = 0; j < nrhs; ++j) std::swap(B[(size_t)(c + j * k)], B[(size_t)(piv + j * k)]);
        }
        const double d = 1.0 / A[(size_t)(c + c * k)];
        for (int r = c + 1; r < k; ++r) {
            const double f = A[(size_t)(r + c * k)] * d;
            if (f == 0.0) continue;
            for (int j = c; j < k; ++j) A[(size_t)(r + j * k)] -= f * A[(size_t)(c + j * k)];
            for (int j = 0; j < nrhs; ++j) B[(size_t)(r + j * k)] -= f * B[(size_t)(c + j * k)];
        }
    }
    for (int j = 0; j < nrhs; ++j)
        for (int r = k - 1; r >= 0; --r) {
            double v = B[(size_t)(r + j * k)];
            for (int c = r + 1; c < k; ++c) v -= A[(size_t)(r + c * k)] * B[(size_t)(c + j * k)];
            B[(size_t)(r + j * k)] = v / A[(size_t)(r + r * k)];
        }
    return true;
}
}  // namespace detail

// lb / ub: n entries each (+-infinity for "no bound").  x: in = start (projected onto the box), out = solution.
inline int lbfgsb_minimize(const Objective &fn, int64_t n, double *x, const double *lb, const double *ub, const LbfgsbOptions &o,
                           LbfgsbReport *rep) {
    using namespace detail;
    const double inf = std::numeric_limits<double>::infinity();
    const int mmax = std::max(1, o.m);
    std::vector<std::vector<double>> S, Y;                 // correction pairs, oldest first
    std::vector<double> g((size_t)n), gn((size_t)n), xn((size_t)n), xcp((size_t)n), d((size_t)n), t((size_t)n), xbar((size_t)n);
    std::vector<double> M;                                  // (2 col) x (2 col), column-major: inverse of [[-D, L'], [L, theta S'S]]
    double theta = 1.0;
    int col = 0;
    LbfgsbReport r;
    for (int64_t i = 0; i < n; ++i) x[i] = std::min(std::max(x[i], lb[i]), ub[i]);
    double f = 0;
    int st = fn(x, &f, g.data());
    r.f_calls = 1;
    if (st) return st;
    auto proj_grad_norm = [&](const double *xx, const double *gg) {
        double m = 0;
        for (int64_t i = 0; i < n; ++i) {
            double gi = gg[i];
            if (gi < 0) gi = std::max(xx[i] - ub[i], gi); else gi = std::min(xx[i] - lb[i], gi);
            const double a = std::fabs(gi);
            if (!(a <= m)) m = a;
        }
        return m;
    };
    r.f = f; r.pg_norm = proj_grad_norm(x, g.data());
    if (!std::isfinite(f) || !std::isfinite(r.pg_norm)) { r.status = 4; *rep = r; return 0; }
    auto wrow = [&](int64_t b, double *w) {                 // row b of W = [Y, theta S]
        for (int j = 0; j < col; ++j) { w[j] = Y[(size_t)j][(size_t)b]; w[col + j] = theta * S[(size_t)j][(size_t)b]; }
    };
    auto Wt_times = [&](const double *v, const std::vector<int64_t> *subset, double *out) {   // out = W' v (over a subset of rows)
        for (int j = 0; j < 2 * col; ++j) out[j] = 0.0;
        auto acc = [&](int64_t i) {
            const double vi = v[i];
            if (vi == 0.0) return;
            for (int j = 0; j < col; ++j) { out[j] += Y[(size_t)j][(size_t)i] * vi; out[col + j] += theta * S[(size_t)j][(size_t)i] * vi; }
        };
        if (subset) for (int64_t i : *subset) acc(i); else for (int64_t i = 0; i < n; ++i) acc(i);
    };
    auto Mv = [&](const double *v, double *out) {
        const int k = 2 * col;
        for (int i = 0; i < k; ++i) { double s = 0; for (int j = 0; j < k; ++j) s += M[(size_t)(i + j * k)] * v[j]; out[i] = s; }
    };
    auto rebuild_M = [&]() -> bool {
        const int k = 2 * col;
        std::vector<double> K((size_t)k * k, 0.0);
        for (int i = 0; i < col; ++i)
            for (int j = 0; j < col; ++j) {
                const double sy = dot(S[(size_t)i].data(), Y[(size_t)j].data(), n);
                if (i == j) K[(size_t)(i + j * k)] = -sy;                                         // -D
                if (i > j) { K[(size_t)((col + i) + j * k)] = sy; K[(size_t)(j + (col + i) * k)] = sy; }   // L and L'
                K[(size_t)((col + i) + (col + j) * k)] = theta * dot(S[(size_t)i].data(), S[(size_t)j].data(), n);
            }
        M.assign((size_t)k * k, 0.0);
        for (int i = 0; i < k; ++i) M[(size_t)(i + i * k)] = 1.0;
        return solve_small(K, k, M, k);
    };
    BfgsOptions ls;
    ls.c1 = 1e-3; ls.c2 = 0.9; ls.max_linesearch = 20;       // ftol, gtol and the 20-evaluation cap of lnsrlb
    std::vector<double> p, c, w, tmp, tmp2, v;
    std::vector<int64_t> order, freeset;
    bool restarted = false;
    while (true) {
        if (r.pg_norm <= o.pgtol) { r.status = 0; break; }
        if (r.iterations >= o.maxiter || r.f_calls >= o.maxfun) { r.status = 2; break; }
        const int k = 2 * col;
        p.assign((size_t)k, 0.0); c.assign((size_t)k, 0.0); w.assign((size_t)k, 0.0); tmp.assign((size_t)k, 0.0); tmp2.assign((size_t)k, 0.0);
        // ---- generalized Cauchy point (algorithm CP) ----
        order.clear();
        for (int64_t i = 0; i < n; ++i) {
            const double gi = g[(size_t)i];
            double ti = inf;
            if (gi < 0 && ub[i] < inf) ti = (x[i] - ub[i]) / gi;
            else if (gi > 0 && lb[i] > -inf) ti = (x[i] - lb[i]) / gi;
            t[(size_t)i] = ti;
            d[(size_t)i] = ti == 0.0 ? 0.0 : -gi;
            xcp[(size_t)i] = x[i];
            if (ti > 0 && ti < inf) order.push_back(i);
        }
        std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return t[(size_t)a] < t[(size_t)b] || (t[(size_t)a] == t[(size_t)b] && a < b); });
        if (col > 0) Wt_times(d.data(), nullptr, p.data());
        double fp = -dot(d.data(), d.data(), n);
        double fpp = -theta * fp;
        if (col > 0) { Mv(p.data(), tmp.data()); double s = 0; for (int j = 0; j < k; ++j) s += p[(size_t)j] * tmp[(size_t)j]; fpp -= s; }
        double dtmin = fpp > 0 ? -fp / fpp : inf, told = 0.0;
        size_t nb = 0;
        bool hit_all = false;
        if (fp == 0.0) dtmin = 0.0;                          // projected gradient direction is zero
        while (nb < order.size()) {
            const int64_t b = order[nb];
            const double tb = t[(size_t)b], dt = tb - told;
            if (dtmin < dt) break;
            const double gb = g[(size_t)b];
            xcp[(size_t)b] = d[(size_t)b] > 0 ? ub[b] : lb[b];
            const double zb = xcp[(size_t)b] - x[b];
            fp += dt * fpp + gb * gb + theta * gb * zb;
            fpp -= theta * gb * gb;
            if (col > 0) {
                for (int j = 0; j < k; ++j) c[(size_t)j] += dt * p[(size_t)j];
                wrow(b, w.data());
                Mv(c.data(), tmp.data());
                double wMc = 0, wMp = 0, wMw = 0;
                for (int j = 0; j < k; ++j) wMc += w[(size_t)j] * tmp[(size_t)j];
                Mv(p.data(), tmp.data());
                for (int j = 0; j < k; ++j) wMp += w[(size_t)j] * tmp[(size_t)j];
                Mv(w.data(), tmp.data());
                for (int j = 0; j < k; ++j) wMw += w[(size_t)j] * tmp[(size_t)j];
                fp -= gb * wMc;
                fpp -= 2.0 * gb * wMp + gb * gb * wMw;
                for (int j = 0; j < k; ++j) p[(size_t)j] += gb * w[(size_t)j];
            }
            fpp = std::max(fpp, std::numeric_limits<double>::epsilon() * std::fabs(fpp) + 1e-300);
            d[(size_t)b] = 0.0;
            dtmin = -fp / fpp;
            told = tb;
            ++nb;
            if (nb == order.size() && !(std::any_of(d.begin(), d.end(), [](double q) { return q != 0.0; }))) hit_all = true;
        }
        if (!hit_all) {
            dtmin = std::max(dtmin, 0.0);
            if (!(dtmin < inf)) dtmin = 0.0;
            told += dtmin;
            for (int64_t i = 0; i < n; ++i) if (d[(size_t)i] != 0.0) xcp[(size_t)i] = x[i] + told * d[(size_t)i];
            for (int j = 0; j < k; ++j) c[(size_t)j] += dtmin * p[(size_t)j];
        }
        // ---- subspace minimisation over the variables free at the Cauchy point (direct primal method) ----
        xbar = xcp;
        freeset.clear();
        for (int64_t i = 0; i < n; ++i) if (xcp[(size_t)i] > lb[i] && xcp[(size_t)i] < ub[i]) freeset.push_back(i);
        if (col > 0 && !freeset.empty()) {
            // reduced gradient of the quadratic model at xcp: r = g + theta (xcp - x) - W M c   (free components)
            Mv(c.data(), tmp.data());
            std::vector<double> rc(freeset.size());
            for (size_t q = 0; q < freeset.size(); ++q) {
                const int64_t i = freeset[q];
                wrow(i, w.data());
                double wm = 0;
                for (int j = 0; j < k; ++j) wm += w[(size_t)j] * tmp[(size_t)j];
                rc[q] = g[(size_t)i] + theta * (xcp[(size_t)i] - x[i]) - wm;
            }
            // v = M W_F' r ;  N = I - M (W_F' W_F) / theta ;  v = N^-1 v ;  du = -r/theta - W_F v / theta^2
            v.assign((size_t)k, 0.0);
            std::vector<double> WtW((size_t)k * k, 0.0);
            for (size_t q = 0; q < freeset.size(); ++q) {
                wrow(freeset[q], w.data());
                for (int a = 0; a < k; ++a) {
                    v[(size_t)a] += w[(size_t)a] * rc[q];
                    for (int b2 = 0; b2 < k; ++b2) WtW[(size_t)(a + b2 * k)] += w[(size_t)a] * w[(size_t)b2];
                }
            }
            Mv(v.data(), tmp.data());
            std::vector<double> N((size_t)k * k, 0.0), rhs(tmp.begin(), tmp.begin() + k);
            for (int a = 0; a < k; ++a)
                for (int b2 = 0; b2 < k; ++b2) {
                    double sacc = 0;
                    for (int e = 0; e < k; ++e) sacc += M[(size_t)(a + e * k)] * WtW[(size_t)(e + b2 * k)];
                    N[(size_t)(a + b2 * k)] = (a == b2 ? 1.0 : 0.0) - sacc / theta;
                }
            if (solve_small(N, k, rhs, 1)) {
                std::vector<double> du(freeset.size());
                for (size_t q = 0; q < freeset.size(); ++q) {
                    wrow(freeset[q], w.data());
                    double wv = 0;
                    for (int j = 0; j < k; ++j) wv += w[(size_t)j] * rhs[(size_t)j];
                    du[q] = -rc[q] / theta - wv / (theta * theta);
                }
                // Morales & Nocedal: project xcp + du onto the box; keep it if the step from x is a descent direction,
                // otherwise fall back to the largest feasible fraction of du
                double dd = 0;
                for (size_t q = 0; q < freeset.size(); ++q) {
                    const int64_t i = freeset[q];
                    xbar[(size_t)i] = std::min(std::max(xcp[(size_t)i] + du[q], lb[i]), ub[i]);
                }
                for (int64_t i = 0; i < n; ++i) dd += (xbar[(size_t)i] - x[i]) * g[(size_t)i];
                if (!(dd < 0)) {
                    double alpha = 1.0;
                    for (size_t q = 0; q < freeset.size(); ++q) {
                        const int64_t i = freeset[q];
                        if (du[q] > 0 && ub[i] < inf) alpha = std::min(alpha, (ub[i] - xcp[(size_t)i]) / du[q]);
                        if (du[q] < 0 && lb[i] > -inf) alpha = std::min(alpha, (lb[i] - xcp[(size_t)i]) / du[q]);
                    }
                    for (size_t q = 0; q < freeset.size(); ++q) xbar[(size_t)freeset[q]] = xcp[(size_t)freeset[q]] + alpha * du[q];
                }
            }
        }
        // ---- line search along d = xbar - x, limited to the feasible segment ----
        for (int64_t i = 0; i < n; ++i) d[(size_t)i] = xbar[(size_t)i] - x[i];
        const double dphi0 = dot(g.data(), d.data(), n), dnorm = std::sqrt(dot(d.data(), d.data(), n));
        bool ls_failed = !(dphi0 < 0) || dnorm == 0.0;
        double fnew = f, alpha = 0;
        if (!ls_failed) {
            double stpmx = 1e10;
            bool constrained = false;
            for (int64_t i = 0; i < n; ++i) {
                if (lb[i] > -inf || ub[i] < inf) constrained = true;
                const double di = d[(size_t)i];
                if (di > 0 && ub[i] < inf) stpmx = std::min(stpmx, (ub[i] - x[i]) / di);
                if (di < 0 && lb[i] > -inf) stpmx = std::min(stpmx, (lb[i] - x[i]) / di);
            }
            if (r.iterations == 0 && !constrained) stpmx = 1e10;
            if (constrained && r.iterations > 0) stpmx = std::max(stpmx, 1.0);   // xbar is feasible, so the unit step always is
            const double stp0 = (r.iterations == 0 && col == 0) ? std::min(1.0 / dnorm, stpmx) : 1.0;
            st = wolfe_search(fn, n, x, d.data(), f, dphi0, stp0, ls, xn.data(), gn.data(), &fnew, &alpha, &r.f_calls, stpmx);
            if (st > 0) return st;
            ls_failed = st == -1;
        }
        if (ls_failed) {
            if (col == 0 || restarted) {                      // steepest descent from a clean memory failed too: give up
                if (col == 0) { r.status = 3; break; }
            }
            S.clear(); Y.clear(); col = 0; theta = 1.0; restarted = true;   // refresh the memory and restart (as lnsrlb's info != 0 path)
            continue;
        }
        restarted = false;
        // ---- accept, update the correction pairs ----
        std::vector<double> s_((size_t)n), y_((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            const double xi = std::min(std::max(xn[(size_t)i], lb[i]), ub[i]);   // guard against rounding outside the box
            s_[(size_t)i] = xi - x[i]; y_[(size_t)i] = gn[(size_t)i] - g[(size_t)i]; x[i] = xi;
        }
        g.swap(gn);
        const double fold = f;
        f = fnew;
        ++r.iterations;
        r.f = f; r.pg_norm = proj_grad_norm(x, g.data());
        if (r.pg_norm <= o.pgtol) { r.status = 0; break; }
        if ((fold - f) <= o.factr * std::numeric_limits<double>::epsilon() * std::max(std::max(std::fabs(fold), std::fabs(f)), 1.0)) { r.status = 1; break; }
        const double sy = dot(s_.data(), y_.data(), n), yy = dot(y_.data(), y_.data(), n);
        if (sy > std::numeric_limits<double>::epsilon() * (-dphi0 * alpha) && yy > 0) {
            if (col == mmax) { S.erase(S.begin()); Y.erase(Y.begin()); --col; }
            S.push_back(std::move(s_)); Y.push_back(std::move(y_)); ++col;
            theta = yy / sy;
            if (!rebuild_M()) { S.clear(); Y.clear(); col = 0; theta = 1.0; }   // singular middle matrix: refresh the memory
        }
    }
    *rep = r;
    return 0;
}

// The objective fit_sfh hands to Optim: LogDensityProblems.logdensity_and_gradient(::HierarchicalOptimizer, xvec)
// (fitting/hierarchical/generic_fitting.jl:90-199) negated as fg_map! / fg_mle! do (:306-325), around an `inner` hierarchical
// fg! over the NATURAL variables [R_1..R_nj, all npar model parameters] (mzr.jl:84-215 / amr.jl:78-173) that returns -logL and
// its gradient.  xvec = [log R_j, transformed FREE parameters]; fixed parameters take params0 (:134-136); transforms are
// 1 (log) or 0 (none) -- the reference's -1 branch is unvalidated (:155-159) and refused by the callers.
inline Objective hier_objective(Objective inner, int64_t nj, int npar, const double *params0, const int32_t *transforms,
                                const uint8_t *free_mask, bool jacobian_corrections) {
    std::vector<double> p0(params0, params0 + npar);
    std::vector<int32_t> tf(transforms, transforms + npar);
    std::vector<uint8_t> fr(free_mask, free_mask + npar);
    auto x = std::make_shared<std::vector<double>>((size_t)nj + npar);
    auto G2 = std::make_shared<std::vector<double>>((size_t)nj + npar);
    return [=](const double *xv, double *f, double *g) -> int {
        std::vector<double> &X = *x, &G = *G2;
        for (int64_t i = 0; i < nj; ++i) X[(size_t)i] = std::exp(xv[i]);                                  // :127
        for (int k = 0, q = 0; k < npar; ++k) {
            if (!fr[(size_t)k]) { X[(size_t)nj + k] = p0[(size_t)k]; continue; }                           // :134-136
            const double v = xv[nj + q++];
            X[(size_t)nj + k] = tf[(size_t)k] == 1 ? std::exp(v) : v;                                      // :129-131
        }
        const int st = inner(X.data(), f, G.data());                                                     // :140
        if (st) return st;
        for (int64_t i = 0; i < nj; ++i) {                     // every R_j is log-transformed (:143-145: eachindex(x)[begin:Nbins])
            if (jacobian_corrections) { *f -= xv[i]; g[i] = G[(size_t)i] * X[(size_t)i] - 1.0; }           // :150-154 (log x_i = xvec_i)
            else g[i] = G[(size_t)i] * X[(size_t)i];                                                      // :163-164
        }
        for (int k = 0, q = 0; k < npar; ++k) {
            if (!fr[(size_t)k]) continue;                                                                // :181-189: free ones only
            double gk = G[(size_t)nj + k];
            if (tf[(size_t)k] == 1) {
                const double xk = X[(size_t)nj + k];
                gk *= xk;
                if (jacobian_corrections) { *f -= std::log(xk); gk -= 1.0; }                              // :150-154
            }
            g[nj + q++] = gk;
        }
        return 0;
    };
}

}  // namespace drivers
}  // namespace sfh
#endif  // SFH_DRIVERS_H
