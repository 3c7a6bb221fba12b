// native_host_sanitize.cpp -- the library's host-only C++ (csrc/sfh_drivers.h, csrc/sfh_nuts.h, csrc/sfh_file.h) compiled by
// itself with g++ under AddressSanitizer + UndefinedBehaviorSanitizer and under ThreadSanitizer, and driven the way
// libsfhcuda.so drives it -- with a CPU Poisson objective standing in for the device evaluations.  Test infrastructure
// (tests/test_native_sanitizers.py builds and runs it); nothing here ships.
//
// What is exercised:
//   * bfgs_minimize on theta = log x of a Poisson likelihood (fit_templates' objective, solvers.jl:180-196), and on a coupled convex
//     objective with n = 512 so that for_columns splits the inverse-Hessian passes over host threads -- also through the
//     HessianBackend interface (what the device-resident inverse Hessian implements), which must give identical iterates;
//   * lbfgsb_minimize with lb = 0 (fit_templates_lbfgsb, solvers.jl:82-90) on a problem whose solution has active bounds;
//   * run_chains: 24 chain threads of ragged length parked on one batched log-density (dense mass matrix), a run whose batch
//     function fails half-way (every thread must be torn down), and a run whose chains are empty;
//   * the container: Writer -> commit -> Reader round trip, multi-threaded checksums, a flipped payload byte and a truncated file;
//   * completion by packets (csrc/sfh_packets.h, what sfh_eval_fg polls instead of synchronising the stream): a thread standing in
//     for the finalize kernel delivers every evaluation's packets in random order, each packet as two separately written 8-byte
//     halves in either order, over the stale packets of the previous evaluation -- the waiting side must return exactly this
//     evaluation's values; an evaluation that ends without delivering, a half-delivered packet and a stream error must fail.
// Exit status 0 = every check passed and no sanitizer report (reports abort the process).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <atomic>
#include <thread>

#include "sfh_drivers.h"
#include "sfh_file.h"
#include "sfh_nuts.h"
#include "sfh_packets.h"

static int failures = 0;
#define CHECK(cond)                                                             \
    do {                                                                        \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static double urand() {   // splitmix64 -> [0, 1)
    uint64_t z = (lcg_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// -logL and its gradient with respect to the coefficients: the arithmetic of fitting_base.jl:84-96, :265-285 (CPU stand-in)
struct Poisson {
    int64_t nb, nt;
    std::vector<double> M, data;   // M column-major nb x nt
    Poisson(int64_t nb_, int64_t nt_, int nzero) : nb(nb_), nt(nt_), M((size_t)(nb_ * nt_)), data((size_t)nb_) {
        for (double &v : M) v = urand();
        std::vector<double> x((size_t)nt);
        for (int64_t j = 0; j < nt; ++j) x[(size_t)j] = (j < nzero) ? 0.0 : 1.0 + 9.0 * urand();
        for (int64_t i = 0; i < nb; ++i) {
            double m = 0;
            for (int64_t j = 0; j < nt; ++j) m += M[(size_t)(i + j * nb)] * x[(size_t)j];
            data[(size_t)i] = std::floor(m + 0.5 * std::sqrt(m) * (2 * urand() - 1) + 0.5);   // integer counts near m
        }
    }
    int fg(const double *x, double *f, double *g) const {
        std::vector<double> r((size_t)nb);
        double logl = 0;
        for (int64_t i = 0; i < nb; ++i) {
            double m = 0;
            for (int64_t j = 0; j < nt; ++j) m += M[(size_t)(i + j * nb)] * x[j];
            m = std::max(m, 2.220446049250313e-16);
            const double n = data[(size_t)i];
            logl += n > 0 ? n - m - n * std::log(n / m) : -m;
            r[(size_t)i] = 1.0 - n / m;
        }
        *f = -logl;
        if (g)
            for (int64_t j = 0; j < nt; ++j) {
                double s = 0;
                for (int64_t i = 0; i < nb; ++i) s += M[(size_t)(i + j * nb)] * r[(size_t)i];
                g[j] = s;
            }
        return 0;
    }
};

static void test_bfgs(int64_t nb, int64_t nt, double tol) {
    using namespace sfh::drivers;
    Poisson P(nb, nt, 0);
    std::vector<double> x((size_t)nt);
    Objective obj = [&](const double *th, double *f, double *g) -> int {   // theta = log x, no prior (the "MLE" of fit_templates)
        for (int64_t j = 0; j < nt; ++j) x[(size_t)j] = std::exp(th[j]);
        P.fg(x.data(), f, g);
        for (int64_t j = 0; j < nt; ++j) g[j] *= x[(size_t)j];
        return 0;
    };
    for (int alphaguess = 0; alphaguess < 2; ++alphaguess) {
        std::vector<double> th((size_t)nt, 0.0), invH((size_t)(nt * nt));
        BfgsOptions o;
        o.g_abstol = tol; o.alphaguess = alphaguess; o.maxiter = 3000;
        BfgsReport rep;
        const int st = bfgs_minimize(obj, nt, th.data(), o, &rep, invH.data());
        CHECK(st == 0);
        CHECK(rep.converged == 1 && rep.g_norm <= tol);
        double asym = 0, dmin = 1e300, hmax = 0;
        for (int64_t j = 0; j < nt; ++j) {
            dmin = std::min(dmin, invH[(size_t)(j + j * nt)]);
            for (int64_t i = 0; i <= j; ++i) {
                hmax = std::max(hmax, std::fabs(invH[(size_t)(i + j * nt)]));
                asym = std::max(asym, std::fabs(invH[(size_t)(i + j * nt)] - invH[(size_t)(j + i * nt)]));
            }
        }
        CHECK(dmin > 0 && asym <= 1e-9 * hmax);   // symmetric up to the rounding of the two orders of the rank-two update
        std::printf("bfgs n=%lld alphaguess=%d: %lld iterations, %lld evaluations, |g| = %.2e\n", (long long)nt, alphaguess,
                    (long long)rep.iterations, (long long)rep.f_calls, rep.g_norm);
    }
    // an objective that fails must hand its status back
    Objective bad = [&](const double *, double *, double *) -> int { return 7; };
    std::vector<double> th((size_t)nt, 0.0), invH((size_t)(nt * nt));
    BfgsReport rep;
    CHECK(bfgs_minimize(bad, nt, th.data(), BfgsOptions{}, &rep, invH.data()) == 7);
}

// n = 512 (n * n >= 2^18): for_columns splits the two inverse-Hessian passes of every iteration over host threads.
// Objective: sum_j (exp(theta_j) - a_j theta_j) + (c / 2) (sum_j theta_j)^2 / n -- convex, coupled, minimum known only numerically.
static void test_bfgs_threaded_hessian() {
    using namespace sfh::drivers;
    const int64_t n = 512;
    std::vector<double> a((size_t)n);
    for (double &v : a) v = 0.5 + 4.0 * urand();
    Objective obj = [&](const double *th, double *f, double *g) -> int {
        double s = 0, tot = 0;
        for (int64_t j = 0; j < n; ++j) { s += std::exp(th[j]) - a[(size_t)j] * th[j]; tot += th[j]; }
        *f = s + 0.5 * 3.0 * tot * tot / (double)n;
        for (int64_t j = 0; j < n; ++j) g[j] = std::exp(th[j]) - a[(size_t)j] + 3.0 * tot / (double)n;
        return 0;
    };
    std::vector<double> th((size_t)n, 0.0), invH((size_t)(n * n));
    BfgsOptions o;
    o.g_abstol = 1e-8;
    BfgsReport rep;
    CHECK(bfgs_minimize(obj, n, th.data(), o, &rep, invH.data()) == 0);
    CHECK(rep.converged == 1 && rep.g_norm <= 1e-8);
    // the same run with a backend that keeps the matrix elsewhere (the HessianBackend interface the device-resident variant
    // implements) must follow the same iterates bit for bit
    struct HostBackend : HessianBackend {
        int64_t n;
        std::vector<double> H;
        explicit HostBackend(int64_t n_) : n(n_), H((size_t)(n_ * n_)) {}
        int reset_identity() override { std::fill(H.begin(), H.end(), 0.0); for (int64_t j = 0; j < n; ++j) H[(size_t)(j + j * n)] = 1.0; return 0; }
        int matvec(const double *g, double *q) override { for (int64_t j = 0; j < n; ++j) q[j] = detail::dot(&H[(size_t)(j * n)], g, n); return 0; }
        int rank2(const double *s, const double *Hy, double rho, double cs) override {
            for (int64_t j = 0; j < n; ++j) {
                const double aa = cs * s[j] - rho * Hy[j], bb = -rho * s[j];
                for (int64_t i = 0; i < n; ++i) H[(size_t)(i + j * n)] += aa * s[i] + bb * Hy[i];
            }
            return 0;
        }
        int download(double *out) override { std::copy(H.begin(), H.end(), out); return 0; }
    } hb(n);
    std::vector<double> th2((size_t)n, 0.0), invH2((size_t)(n * n));
    BfgsReport rep2;
    CHECK(bfgs_minimize(obj, n, th2.data(), o, &rep2, invH2.data(), &hb) == 0);
    CHECK(rep2.iterations == rep.iterations && rep2.f_calls == rep.f_calls && rep2.f == rep.f);
    CHECK(memcmp(th.data(), th2.data(), (size_t)n * 8) == 0 && memcmp(invH.data(), invH2.data(), (size_t)(n * n) * 8) == 0);
    std::printf("bfgs n=512 (threaded inverse Hessian): %lld iterations, %lld evaluations, |g| = %.2e; backend run identical\n",
                (long long)rep.iterations, (long long)rep.f_calls, rep.g_norm);
}

static void test_lbfgsb() {
    using namespace sfh::drivers;
    const int64_t nb = 400, nt = 40;
    Poisson P(nb, nt, 6);   // six true coefficients are exactly zero: bounds become active
    Objective obj = [&](const double *x, double *f, double *g) -> int { return P.fg(x, f, g); };
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> x((size_t)nt, 1.0), lb((size_t)nt, 0.0), ub((size_t)nt, inf), g((size_t)nt);
    LbfgsbOptions o;
    LbfgsbReport rep;
    const int st = lbfgsb_minimize(obj, nt, x.data(), lb.data(), ub.data(), o, &rep);
    CHECK(st == 0);
    CHECK(rep.status == 0 || rep.status == 1);
    double f = 0, kkt = 0;
    P.fg(x.data(), &f, g.data());
    int nactive = 0;
    for (int64_t j = 0; j < nt; ++j) {
        CHECK(x[(size_t)j] >= 0);
        if (x[(size_t)j] == 0) { ++nactive; CHECK(g[(size_t)j] >= -1e-5); }   // at a bound the gradient points outward
        else kkt = std::max(kkt, std::fabs(g[(size_t)j]));
    }
    CHECK(kkt < 1e-3);
    std::printf("lbfgsb: status %d, %lld iterations, %lld evaluations, %d active bounds, free |g| = %.2e\n", rep.status,
                (long long)rep.iterations, (long long)rep.f_calls, nactive, kkt);
    // box with finite upper bounds that cut the solution off
    std::vector<double> x2((size_t)nt, 1.0), ub2((size_t)nt, 3.0);
    CHECK(lbfgsb_minimize(obj, nt, x2.data(), lb.data(), ub2.data(), o, &rep) == 0);
    for (int64_t j = 0; j < nt; ++j) CHECK(x2[(size_t)j] >= 0 && x2[(size_t)j] <= 3.0);
}

static void test_nuts() {
    using namespace sfh::nuts;
    const int64_t n = 6, nchains = 24;
    // target N(mu, Sigma), Sigma = A A' + I; log-density and gradient for C chains at once
    std::vector<double> Sigma((size_t)(n * n), 0.0), Prec((size_t)(n * n), 0.0), mu((size_t)n);
    {
        std::vector<double> A((size_t)(n * n));
        for (double &v : A) v = urand() - 0.5;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) {
                double s = (i == j) ? 1.0 : 0.0;
                for (int64_t k = 0; k < n; ++k) s += A[(size_t)(i + k * n)] * A[(size_t)(j + k * n)];
                Sigma[(size_t)(i + j * n)] = s;
            }
        for (int64_t i = 0; i < n; ++i) mu[(size_t)i] = 3.0 * urand();
        // precision by Gauss-Jordan on [Sigma | I]
        std::vector<double> W = Sigma;
        for (int64_t i = 0; i < n; ++i) Prec[(size_t)(i + i * n)] = 1.0;
        for (int64_t c = 0; c < n; ++c) {
            const double p = W[(size_t)(c + c * n)];
            for (int64_t j = 0; j < n; ++j) { W[(size_t)(c + j * n)] /= p; Prec[(size_t)(c + j * n)] /= p; }
            for (int64_t r = 0; r < n; ++r) {
                if (r == c) continue;
                const double fct = W[(size_t)(r + c * n)];
                for (int64_t j = 0; j < n; ++j) { W[(size_t)(r + j * n)] -= fct * W[(size_t)(c + j * n)]; Prec[(size_t)(r + j * n)] -= fct * Prec[(size_t)(c + j * n)]; }
            }
        }
    }
    int64_t calls = 0, fail_after = -1;
    BatchLogDensity fn = [&](const double *Th, int64_t C, double *lp, double *grad) -> int {
        if (fail_after >= 0 && calls >= fail_after) return 42;
        ++calls;
        for (int64_t c = 0; c < C; ++c) {
            double q = 0;
            for (int64_t i = 0; i < n; ++i) {
                double s = 0;
                for (int64_t j = 0; j < n; ++j) s += Prec[(size_t)(i + j * n)] * (Th[c * n + j] - mu[(size_t)j]);
                grad[c * n + i] = -s;
                q += s * (Th[c * n + i] - mu[(size_t)i]);
            }
            lp[c] = -0.5 * q;
        }
        return 0;
    };
    std::vector<int64_t> nsteps((size_t)nchains);
    int64_t total = 0;
    for (int64_t c = 0; c < nchains; ++c) { nsteps[(size_t)c] = 40 + 7 * (c % 5); total += nsteps[(size_t)c]; }   // ragged
    std::vector<double> theta0((size_t)(n * nchains)), samples((size_t)(n * total)), logps((size_t)total), eps((size_t)nchains);
    for (double &v : theta0) v = urand();
    for (int kind = 0; kind < 3; ++kind) {
        Mass mass;
        std::vector<double> diag((size_t)n);
        for (int64_t i = 0; i < n; ++i) diag[(size_t)i] = Sigma[(size_t)(i + i * n)];
        CHECK(mass.init(kind, n, kind == 1 ? diag.data() : Sigma.data()));
        Options o;
        o.max_depth = 6; o.nwarmup = 60; o.seed = 1234 + (uint64_t)kind;
        Stats stats;
        calls = 0; fail_after = -1;
        const int st = run_chains(fn, n, nchains, theta0.data(), nsteps.data(), mass, o, samples.data(), logps.data(), eps.data(), &stats);
        CHECK(st == 0);
        CHECK(stats.n_batches == calls && stats.n_evals >= total);
        double err = 0;
        for (int64_t i = 0; i < n; ++i) {
            double m = 0;
            for (int64_t t = 0; t < total; ++t) m += samples[(size_t)(i + t * n)];
            err = std::max(err, std::fabs(m / (double)total - mu[(size_t)i]) / std::sqrt(Sigma[(size_t)(i + i * n)]));
        }
        CHECK(err < 0.35);   // ~1200 correlated draws: the mean within a third of a standard deviation
        for (int64_t c = 0; c < nchains; ++c) CHECK(eps[(size_t)c] > 0 && std::isfinite(eps[(size_t)c]));
        std::printf("nuts mass kind %d: %lld batches, %lld evaluations, worst mean error %.3f sigma\n", kind, (long long)stats.n_batches,
                    (long long)stats.n_evals, err);
    }
    {   // the batch function fails after 25 batches: the status comes back and all 24 threads are joined
        Mass mass;
        CHECK(mass.init(0, n, nullptr));
        Options o;
        o.max_depth = 5; o.nwarmup = 20; o.seed = 99;
        calls = 0; fail_after = 25;
        CHECK(run_chains(fn, n, nchains, theta0.data(), nsteps.data(), mass, o, samples.data(), logps.data(), eps.data(), nullptr) == 42);
    }
    {   // chains that are asked for no draws at all (ragged down to zero) next to live ones
        Mass mass;
        CHECK(mass.init(0, n, nullptr));
        Options o;
        o.max_depth = 5; o.nwarmup = 10; o.seed = 5;
        std::vector<int64_t> ns((size_t)nchains, 0);
        ns[3] = 5; ns[17] = 2;
        calls = 0; fail_after = -1;
        Stats stats;
        CHECK(run_chains(fn, n, nchains, theta0.data(), ns.data(), mass, o, samples.data(), logps.data(), eps.data(), &stats) == 0);
        for (int64_t t = 0; t < 7; ++t) CHECK(std::isfinite(logps[(size_t)t]));
    }
    {   // a non-positive-definite dense mass matrix is refused
        Mass mass;
        std::vector<double> bad((size_t)(n * n), 1.0);
        CHECK(!mass.init(2, n, bad.data()));
    }
}

static void test_file(const char *dir) {
    using namespace sfh::file;
    const std::string path = std::string(dir) + "/sanitize_roundtrip.sfh";
    const int64_t nb = 3000, nt = 41;   // 984 000 bytes of payload: not a multiple of the 4096-byte alignment
    std::vector<double> M((size_t)(nb * nt));
    for (double &v : M) v = urand();
    std::vector<float> small(7);
    for (float &v : small) v = (float)urand();
    std::vector<unsigned char> mask = {1, 0, 1};
    std::vector<ArraySpec> specs(4);
    specs[0].name = "models"; specs[0].dtype = 1; specs[0].ndim = 2; specs[0].dims[0] = nb; specs[0].dims[1] = nt; specs[0].ptr = nullptr;
    specs[1].name = "small"; specs[1].dtype = 0; specs[1].ndim = 1; specs[1].dims[0] = 7; specs[1].ptr = small.data();
    specs[2].name = "free_mask"; specs[2].dtype = 3; specs[2].ndim = 1; specs[2].dims[0] = 3; specs[2].ptr = mask.data();
    specs[3].name = "empty"; specs[3].dtype = 1; specs[3].ndim = 1; specs[3].dims[0] = 0; specs[3].ptr = nullptr;
    const int64_t attrs[8] = {nb, nt, 1, 2, 3, 4, 5, 6};
    std::string err;
    {
        Writer w;
        CHECK(w.begin(path.c_str(), 1, attrs, specs, &err));
        memcpy(w.section(0), M.data(), M.size() * 8);   // a section produced in place through the mapping
        CHECK(w.commit(&err));
    }
    {
        Reader r;
        CHECK(r.open(path.c_str(), &err));
        CHECK(r.count() == 4 && r.header().kind == 1 && r.header().attrs[1] == nt);
        const int i = r.find("models");
        CHECK(i == 0 && r.entry(i).nbytes == (uint64_t)(nb * nt * 8) && r.verify(i));
        CHECK(memcmp(r.data(i), M.data(), M.size() * 8) == 0);
        CHECK(r.verify(r.find("small")) && r.verify(r.find("free_mask")) && r.verify(r.find("empty")) && r.find("nope") == -1);
        CHECK(memcmp(r.data(r.find("free_mask")), mask.data(), 3) == 0);
    }
    CHECK(checksum(M.data(), M.size() * 8, 1) == checksum(M.data(), M.size() * 8, 7));   // independent of the thread count
    CHECK(checksum(M.data(), M.size() * 8 - 3, 1) == checksum(M.data(), M.size() * 8 - 3, 5));   // ragged tail
    {   // flip one payload byte: open still succeeds, verify of that array fails
        FILE *f = fopen(path.c_str(), "r+b");
        CHECK(f != nullptr);
        if (f) {
            fseek(f, 4096 + 12345, SEEK_SET);
            int c = fgetc(f);
            fseek(f, 4096 + 12345, SEEK_SET);
            fputc(c ^ 0x10, f);
            fclose(f);
        }
        Reader r;
        CHECK(r.open(path.c_str(), &err));
        CHECK(!r.verify(0) && r.verify(1));
    }
    {   // truncated file: refused at open
        CHECK(truncate(path.c_str(), 8192) == 0);
        Reader r;
        CHECK(!r.open(path.c_str(), &err));
    }
    {   // an abandoned writer leaves nothing behind; bad specs are refused
        const std::string p2 = std::string(dir) + "/sanitize_abandoned.sfh";
        {
            Writer w;
            CHECK(w.begin(p2.c_str(), 0, nullptr, specs, &err));
        }
        CHECK(access(p2.c_str(), F_OK) != 0);
        std::vector<ArraySpec> dup = {specs[1], specs[1]};
        Writer w;
        CHECK(!w.begin(p2.c_str(), 0, nullptr, dup, &err));
    }
    unlink(path.c_str());
}

// ---- completion by packets ----------------------------------------------------------------------------------------------
static double packet_value(uint32_t epoch, size_t j) { return 1e3 * (double)epoch + (double)j + 0.25; }

static void test_packets() {
    namespace pk = sfh_packets;
    const size_t n = 301, rest_at = 8, npk = rest_at + n;
    std::vector<uint64_t> buf(2 * npk, 0);
    std::atomic<uint32_t> request{0}, delivered{0};
    std::atomic<bool> stop{false};
    // the "device": for every requested epoch, all packets in a random order, halves written separately in a random order
    std::thread dev([&] {
        uint64_t st = 12345;
        auto rnd = [&] { st = st * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(st >> 33); };
        uint32_t done = 0;
        std::vector<size_t> order(n + 1);
        while (!stop.load(std::memory_order_acquire)) {
            const uint32_t want = request.load(std::memory_order_acquire);
            if (want == done) { std::this_thread::yield(); continue; }
            for (size_t i = 0; i <= n; ++i) order[i] = i;
            for (size_t i = n; i > 0; --i) std::swap(order[i], order[rnd() % (i + 1)]);
            for (size_t i = 0; i <= n; ++i) {
                const size_t j = order[i], at = j == 0 ? 0 : rest_at + j - 1;
                uint64_t p[2];
                pk::encode(p, packet_value(want, j), want);
                const int first_half = (int)(rnd() & 1u);
                __atomic_store_n(&buf[2 * at + first_half], p[first_half], __ATOMIC_RELEASE);
                if ((rnd() & 7u) == 0) std::this_thread::yield();            // the other half arrives later
                __atomic_store_n(&buf[2 * at + 1 - first_half], p[1 - first_half], __ATOMIC_RELEASE);
            }
            done = want;
            delivered.store(done, std::memory_order_release);
        }
    });
    std::vector<double> G(n);
    bool all_ok = true;
    for (uint32_t ep = 1; ep <= 400; ++ep) {
        request.store(ep, std::memory_order_release);
        double f = 0.0;
        size_t missing = 0;
        const bool f_only = ep % 3 == 0;   // logL-only call: packet 0 alone is waited for, and the NEXT call starts while this
                                           // epoch's gradient packets are still landing (they must not be taken for the next epoch's)
        const int r = pk::wait(buf.data(), ep, &f, f_only ? nullptr : G.data(), n, rest_at,
                               [&]() -> int { return delivered.load(std::memory_order_acquire) == ep ? pk::kDrained : pk::kRunning; }, &missing, 64);
        all_ok = all_ok && r == pk::kOk && f == packet_value(ep, 0);
        for (size_t j = 1; j <= n && all_ok && !f_only; ++j) all_ok = G[j - 1] == packet_value(ep, j);
        if (!all_ok) { std::printf("packets: epoch %u failed (r = %d, missing %zu)\n", ep, r, missing); break; }
    }
    CHECK(all_ok);
    while (delivered.load(std::memory_order_acquire) != request.load()) std::this_thread::yield();
    stop.store(true, std::memory_order_release);
    dev.join();
    // an evaluation that drains without delivering: must fail with the index of the missing packet, not hang
    {
        double f = 0.0;
        size_t missing = 999;
        int calls = 0;
        const int r = pk::wait(buf.data(), 100000u, &f, G.data(), n, rest_at, [&]() -> int { ++calls; return pk::kDrained; }, &missing, 16);
        CHECK(r == pk::kMissing && missing == 0 && calls == 1);
    }
    // a packet of which only ONE half ever arrives is never accepted
    {
        uint64_t p[2];
        pk::encode(p, 7.5, 500000u);
        __atomic_store_n(&buf[0], p[0], __ATOMIC_RELEASE);
        pk::encode(p, packet_value(500000u, 1), 500000u);
        __atomic_store_n(&buf[2 * rest_at], p[0], __ATOMIC_RELEASE);
        __atomic_store_n(&buf[2 * rest_at + 1], p[1], __ATOMIC_RELEASE);
        double f = 0.0;
        size_t missing = 999;
        const int r = pk::wait(buf.data(), 500000u, &f, G.data(), n, rest_at, [&]() -> int { return pk::kDrained; }, &missing, 16);
        CHECK(r == pk::kMissing && missing == 0);
        double v = 0.0;
        CHECK(!pk::try_read(buf.data(), 500000u, &v) && pk::try_read(buf.data() + 2 * rest_at, 500000u, &v) && v == packet_value(500000u, 1));
    }
    // a stream error ends the wait at once
    {
        double f = 0.0;
        size_t missing = 0;
        const int r = pk::wait(buf.data(), 600000u, &f, nullptr, 0, rest_at, [&]() -> int { return pk::kFailed; }, &missing, 4);
        CHECK(r == pk::kStreamError);
    }
    // encode / try_read round trip over awkward values
    {
        const double vals[] = {0.0, -0.0, 1.0, -1.5e300, 4.9e-324, 1.0 / 0.0, -1.0 / 0.0};
        for (double x : vals) {
            uint64_t p[2];
            double v = 1.0;
            pk::encode(p, x, 0xfffffffeu);
            CHECK(pk::try_read(p, 0xfffffffeu, &v) && std::memcmp(&v, &x, 8) == 0 && !pk::try_read(p, 0xfffffffdu, &v));
        }
    }
}

int main(int argc, char **argv) {
    const char *dir = argc > 1 ? argv[1] : "/tmp";
    test_packets();
    test_bfgs(300, 24, 1e-8);
    test_bfgs_threaded_hessian();
    test_lbfgsb();
    test_nuts();
    test_file(dir);
    std::printf(failures ? "%d CHECKS FAILED\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
