// examples/fit_sfh_native.cpp -- a complete hierarchical fit with nothing but the C-ABI of libsfhcuda.so (no Python, no Julia):
//
//   synthetic template stack generated on the device (BASELINE config 3 shape by default: 200x300 bins x 60 ages x 40 [M/H])
//   -> fit_sfh: MAP then MLE, dense BFGS inside the library          (generic_fitting.jl:242-411  -> sfh_fit_sfh_bfgs)
//   -> tsample_sfh: short NUTS chains, one batched pass per round     (generic_fitting.jl:564-665  -> sfh_sample_sfh_nuts)
//   -> the stack and the results written to container files           (sfh_stack_save, sfh_file_write)
//
// build:  g++ -O2 -std=c++17 -Iinclude examples/fit_sfh_native.cpp -o fit_sfh_native -Lstarformationhistories.jl_b200
//             -l:libsfhcuda.so -Wl,-rpath,$PWD/starformationhistories.jl_b200          (one command line)
// run:    ./fit_sfh_native [nbins_x nbins_y n_ages n_mh [outdir]]        (exits 0 with a message when there is no GPU)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sfhcuda.h"

#define OK(call)                                                                           \
    do {                                                                                   \
        const int st_ = (call);                                                            \
        if (st_ != SFH_OK) {                                                               \
            std::fprintf(stderr, "%s -> status %d: %s\n", #call, st_, sfh_last_error());   \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

// splitmix64 -> U(0,1): the true star-formation history of the synthetic galaxy
static double u01(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return (double)((z ^ (z >> 31)) >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
    const int64_t nx = argc > 2 ? std::atoll(argv[1]) : 200, ny = argc > 2 ? std::atoll(argv[2]) : 300;
    const int64_t nj = argc > 4 ? std::atoll(argv[3]) : 60, nk = argc > 4 ? std::atoll(argv[4]) : 40;
    const std::string outdir = argc > 5 ? argv[5] : ".";
    const int64_t nb = nx * ny, nt = nj * nk;
    int ndev = 0;
    OK(sfh_device_count(&ndev));
    if (ndev == 0) {
        std::printf("no CUDA device: libsfhcuda has no CPU fallback, nothing to run\n");
        return 0;
    }
    // ---- the grid, the truth, and the coefficients the truth implies (calculate_coeffs on the device, mzr.jl:50-79) ----
    std::vector<double> logAge((size_t)nt), MH((size_t)nt), truth((size_t)nj + 3), coeffs((size_t)nt, 1.0);
    for (int64_t j = 0; j < nj; ++j)
        for (int64_t k = 0; k < nk; ++k) {
            logAge[(size_t)(j * nk + k)] = 10.1 - 3.5 * (double)j / (double)(nj > 1 ? nj - 1 : 1);
            MH[(size_t)(j * nk + k)] = -2.5 + 2.5 * (double)k / (double)(nk > 1 ? nk - 1 : 1);
        }
    uint64_t seed = 94823;
    for (int64_t j = 0; j < nj; ++j) truth[(size_t)j] = 1e6 * u01(seed);
    truth[(size_t)nj] = 1.0; truth[(size_t)nj + 1] = -2.0; truth[(size_t)nj + 2] = 0.2;   // PowerLawMZR(1, -2, 6), GaussianDispersion(0.2)
    const double mh_fixed[4] = {6.0, 0, 0, 0};

    // a throw-away stack only to expand the truth into coefficients, then the real one whose data are Poisson(M x_true)
    sfh_stack *stack = nullptr;
    sfh_ctx *ctx = nullptr;
    OK(sfh_stack_create_synthetic(&stack, nb, nt, SFH_F64, 94823, 1e-5, coeffs.data(), nullptr));
    OK(sfh_ctx_create(stack, nullptr, &ctx));
    int64_t n_ages = 0;
    OK(sfh_hier_bind(ctx, logAge.data(), MH.data(), &n_ages));
    OK(sfh_calculate_coeffs(ctx, SFH_MH_POWERLAW_MZR, mh_fixed, SFH_DISP_GAUSSIAN, truth.data(), coeffs.data()));
    OK(sfh_ctx_destroy(ctx));
    OK(sfh_stack_destroy(stack));
    OK(sfh_stack_create_synthetic(&stack, nb, nt, SFH_F64, 94823, 1e-5, coeffs.data(), nullptr));
    OK(sfh_ctx_create(stack, nullptr, &ctx));
    OK(sfh_hier_bind(ctx, logAge.data(), MH.data(), &n_ages));
    std::printf("stack: %lld bins x %lld templates (%lld ages) on the device\n", (long long)nb, (long long)nt, (long long)n_ages);

    // ---- fit_sfh: start 1.5x off in the masses, (1.2, -2.2, 0.25) in the model parameters (mzr_test.jl:178-182) ----
    const double params0[3] = {1.2, -2.2, 0.25};
    const int32_t transforms[3] = {1, 0, 1};   // PowerLawMZR (1, 0), GaussianDispersion (1,)
    const uint8_t free_mask[3] = {1, 1, 1};
    const int64_t n = nj + 3;
    std::vector<double> x((size_t)n), invH_map((size_t)(n * n)), invH_mle((size_t)(n * n));
    // renormalize_x0 (fitting/utilities.jl:104-115): scale the start so that the model's total counts match the data's
    std::vector<double> v0(truth), c0((size_t)nt), comp((size_t)nb), data((size_t)nb);
    for (int64_t j = 0; j < nj; ++j) v0[(size_t)j] = 1.5 * truth[(size_t)j];
    v0[(size_t)nj] = params0[0]; v0[(size_t)nj + 1] = params0[1]; v0[(size_t)nj + 2] = params0[2];
    OK(sfh_calculate_coeffs(ctx, SFH_MH_POWERLAW_MZR, mh_fixed, SFH_DISP_GAUSSIAN, v0.data(), c0.data()));
    OK(sfh_composite(ctx, c0.data(), comp.data()));
    OK(sfh_stack_download(stack, nullptr, data.data()));
    double sum_c = 0, sum_d = 0;
    for (int64_t i = 0; i < nb; ++i) { sum_c += comp[(size_t)i]; sum_d += data[(size_t)i]; }
    for (int64_t j = 0; j < nj; ++j) x[(size_t)j] = std::log(v0[(size_t)j] * (sum_c > 0 ? sum_d / sum_c : 1.0));
    x[(size_t)nj] = std::log(params0[0]); x[(size_t)nj + 1] = params0[1]; x[(size_t)nj + 2] = std::log(params0[2]);

    sfh_bfgs_opts bo;
    std::memset(&bo, 0, sizeof bo);
    bo.struct_size = (int32_t)sizeof bo;
    bo.g_abstol = 1e-6;
    sfh_bfgs_report rmap, rmle;
    OK(sfh_fit_sfh_bfgs(ctx, SFH_MH_POWERLAW_MZR, mh_fixed, SFH_DISP_GAUSSIAN, params0, transforms, free_mask, 1, x.data(), &bo, &rmap, invH_map.data()));
    std::vector<double> x_map(x);
    OK(sfh_fit_sfh_bfgs(ctx, SFH_MH_POWERLAW_MZR, mh_fixed, SFH_DISP_GAUSSIAN, params0, transforms, free_mask, 0, x.data(), &bo, &rmle, invH_mle.data()));
    std::printf("MAP: %lld iterations, %lld evaluations, |g| = %.2e, converged = %d\n", (long long)rmap.iterations, (long long)rmap.f_calls, rmap.g_norm, rmap.converged);
    std::printf("MLE: %lld iterations, %lld evaluations, |g| = %.2e, converged = %d\n", (long long)rmle.iterations, (long long)rmle.f_calls, rmle.g_norm, rmle.converged);
    std::printf("alpha, beta, sigma = %.4f, %.4f, %.4f   (truth 1, -2, 0.2)\n", std::exp(x[(size_t)nj]), x[(size_t)nj + 1], std::exp(x[(size_t)nj + 2]));
    int within = 0;
    for (int64_t j = 0; j < nj; ++j) {   // sigma of R_j by the delta method: sqrt(invH_jj) * R_j  (generic_fitting.jl:352-407)
        const double mu = std::exp(x_map[(size_t)j]), sd = std::sqrt(std::fabs(invH_map[(size_t)(j + j * n)])) * mu;
        within += std::fabs(mu - truth[(size_t)j]) < 3 * sd;
    }
    std::printf("%d of %lld stellar-mass coefficients within 3 sigma of the truth\n", within, (long long)nj);

    // ---- tsample_sfh: 16 chains x 25 draws from the MLE, MAP.invH as M^-1, epsilon = 0.05 (generic_fitting.jl:564-665) ----
    const int64_t nchains = 16, len = 25;
    std::vector<double> starts((size_t)(n * nchains)), samples((size_t)(n * nchains * len)), logps((size_t)(nchains * len)), steps((size_t)nchains);
    std::vector<int64_t> lens((size_t)nchains, len);
    for (int64_t c = 0; c < nchains; ++c)   // (the reference draws the starts from MvNormal(MLE, MAP.invH); the MLE itself keeps this example short)
        for (int64_t i = 0; i < n; ++i) starts[(size_t)(c * n + i)] = x[(size_t)i];
    sfh_nuts_opts no;
    std::memset(&no, 0, sizeof no);
    no.struct_size = (int32_t)sizeof no; no.max_depth = 5; no.eps0 = 0.05; no.seed = 1; no.mass_kind = 2;
    int64_t nbatches = 0, nevals = 0;
    OK(sfh_sample_sfh_nuts(ctx, SFH_MH_POWERLAW_MZR, mh_fixed, SFH_DISP_GAUSSIAN, params0, transforms, free_mask, nchains, starts.data(), lens.data(),
                           invH_map.data(), &no, samples.data(), logps.data(), steps.data(), &nbatches, &nevals));
    std::printf("NUTS: %lld draws, %lld gradient evaluations served in %lld batched passes (%.1f chains per pass)\n", (long long)(nchains * len),
                (long long)nevals, (long long)nbatches, (double)nevals / (double)(nbatches ? nbatches : 1));

    // ---- persist: the stack (device -> memory-mapped file) and the results ----
    OK(sfh_stack_save(stack, (outdir + "/stack.sfh").c_str(), nx, ny, logAge.data(), MH.data()));
    sfh_array_desc d[4];
    std::memset(d, 0, sizeof d);
    const char *names[4] = {"mle/x", "map/x", "map/invH", "posterior_x"};
    const void *ptrs[4] = {x.data(), x_map.data(), invH_map.data(), samples.data()};
    for (int a = 0; a < 4; ++a) { std::strcpy(d[a].name, names[a]); d[a].dtype = SFH_F64; d[a].ndim = 1; d[a].dims[0] = n; }
    d[2].ndim = 2; d[2].dims[1] = n;
    d[3].ndim = 2; d[3].dims[1] = nchains * len;
    OK(sfh_file_write((outdir + "/fit.sfh").c_str(), SFH_FILE_RESULT, nullptr, 4, d, ptrs));
    std::printf("wrote %s/stack.sfh and %s/fit.sfh\n", outdir.c_str(), outdir.c_str());
    OK(sfh_ctx_destroy(ctx));
    OK(sfh_stack_destroy(stack));
    return 0;
}
