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
                        const BfgsOptions &o, double *x_new, double *g_new, double *f_new, double *alpha_out, int64_t *f_calls) {
    using namespace detail;
    int evals = 0;
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
        // still descending steeply: extrapolate to where the secant through the two slopes vanishes, kept inside
        // [1.25, 10] x the current step (the scaled first trial can be orders of magnitude short of the minimiser)
        double next = 10.0 * a1;
        if (da1 > da0) next = std::min(10.0 * a1, std::max(1.25 * a1, a1 - da1 * (a1 - a0) / (da1 - da0)));
        a0 = a1; fa0 = fa1; da0 = da1;
        a1 = next;
    }
    return -1;
}

// Dense BFGS on the inverse Hessian (column-major n x n in invH, which must hold n*n doubles; it is initialised to the
// identity here, as Optim does).  x: in = start, out = minimiser.
inline int bfgs_minimize(const Objective &fn, int64_t n, double *x, const BfgsOptions &o, BfgsReport *rep, double *invH) {
    using namespace detail;
    std::vector<double> g((size_t)n), gn((size_t)n), xn((size_t)n), p((size_t)n), s((size_t)n), y((size_t)n), Hy((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        double *col = invH + j * n;
        std::fill(col, col + n, 0.0);
        col[j] = 1.0;
    }
    BfgsReport r;
    double f = 0;
    int st = fn(x, &f, g.data());
    r.f_calls = 1;
    if (st) return st;
    r.f = f; r.g_norm = infnorm(g.data(), n);
    if (!std::isfinite(f) || !std::isfinite(r.g_norm)) { r.status = 3; *rep = r; return 0; }
    double f_prev = f + std::sqrt(dot(g.data(), g.data(), n)) / 2.0;
    while (true) {
        if (r.g_norm <= o.g_abstol) { r.converged = 1; r.status = 0; break; }
        if (r.iterations >= o.maxiter) { r.status = 1; break; }
        // p = -H g
        for_columns(n, [&](int64_t j0, int64_t j1) {
            for (int64_t j = j0; j < j1; ++j) p[(size_t)j] = -dot(invH + j * n, g.data(), n);   // H symmetric: column j = row j
        });
        double dphi0 = dot(g.data(), p.data(), n);
        if (!(dphi0 < 0)) {   // not a descent direction (H lost positive definiteness to rounding): restart from the identity
            for (int64_t j = 0; j < n; ++j) { double *col = invH + j * n; std::fill(col, col + n, 0.0); col[j] = 1.0; p[(size_t)j] = -g[(size_t)j]; }
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
        const double ys = dot(y.data(), s.data(), n);
        const double rho = ys != 0.0 ? 1.0 / ys : 1000.0;
        if (!(ys > 0)) continue;   // curvature condition violated (cannot happen with a Wolfe step up to rounding): skip the update
        for_columns(n, [&](int64_t j0, int64_t j1) {
            for (int64_t j = j0; j < j1; ++j) Hy[(size_t)j] = dot(invH + j * n, y.data(), n);
        });
        const double yHy = dot(y.data(), Hy.data(), n);
        const double cs = rho * rho * yHy + rho;
        for_columns(n, [&](int64_t j0, int64_t j1) {
            for (int64_t j = j0; j < j1; ++j) {
                double *col = invH + j * n;
                const double a = cs * s[(size_t)j] - rho * Hy[(size_t)j], b = -rho * s[(size_t)j];
                for (int64_t i = 0; i < n; ++i) col[i] += a * s[(size_t)i] + b * Hy[(size_t)i];
            }
        });
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
