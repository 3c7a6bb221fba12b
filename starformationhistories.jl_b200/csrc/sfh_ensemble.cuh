// Device-resident affine-invariant ensemble sampler (Goodman & Weare stretch move) around the batched-walker
// kernel K6.  The reference hands MCMCModel to KissMCMC.emcee (src/fitting/mcmc_sample.jl:97-108), which proposes,
// evaluates and accepts walker by walker on host threads; here the proposal, the log-likelihood of a whole
// half-ensemble (K6) and the accept/reject all stay on the device, so a step costs no host round trip and no
// transfer of the T x W walker matrix.
//
// Random numbers: Philox4x32-10 (sfh_small.cuh), counter = (step, half, walker), stream = draw kind, key = seed,
// so a run is reproducible for any launch geometry / GPU count and can be restated exactly on the host (tests).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sfh_small.cuh"

namespace sfh {

enum : uint32_t { kDrawStretch = 16u, kDrawPartner = 17u, kDrawAccept = 18u };

__host__ __device__ __forceinline__ uint64_t ensemble_counter(int64_t step, int h, int64_t a) {
    return ((uint64_t)step << 33) | ((uint64_t)(h & 1) << 32) | (uint64_t)(uint32_t)a;
}

// One warp per active walker a of half h (walkers [h*half, (h+1)*half)); partner drawn from the other half.
//   z = ((a_scale-1) u + 1)^2 / a_scale        (density ~ 1/sqrt(z) on [1/a_scale, a_scale])
//   P[:, a] = X[:, j] + z (X[:, a] - X[:, j])
__global__ void sfh_stretch_propose_kernel(const double *__restrict__ X, int64_t nt, int64_t half, int h, int64_t step,
                                           uint64_t seed, double a_scale, double *__restrict__ P, double *__restrict__ zout) {
    const int lane = threadIdx.x & 31;
    const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= half) return;
    const uint64_t ctr = ensemble_counter(step, h, a);
    const double u = philox_u01(ctr, seed, kDrawStretch);
    const double t = (a_scale - 1.0) * u + 1.0;
    const double z = t * t / a_scale;
    int64_t j = (int64_t)(philox_u01(ctr, seed, kDrawPartner) * (double)half);
    if (j >= half) j = half - 1;
    const double *xa = X + (size_t)nt * (size_t)(h * half + a);
    const double *xj = X + (size_t)nt * (size_t)((1 - h) * half + j);
    double *p = P + (size_t)nt * (size_t)a;
    for (int64_t k = lane; k < nt; k += 32) {
        const double pj = xj[k];
        p[k] = pj + z * (xa[k] - pj);
    }
    if (lane == 0) zout[a] = z;
}

// accept with probability min(1, z^(T-1) L(P)/L(X)); a proposal with a non-finite log-likelihood (a negative
// coefficient, mcmc_sample.jl:15-19) is never accepted.
__global__ void sfh_stretch_accept_kernel(double *__restrict__ X, double *__restrict__ lp, const double *__restrict__ P,
                                          const double *__restrict__ lpp, const double *__restrict__ z, int64_t nt,
                                          int64_t half, int h, int64_t step, uint64_t seed,
                                          unsigned long long *__restrict__ n_accept) {
    const int lane = threadIdx.x & 31;
    const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= half) return;
    const int64_t w = h * half + a;
    const double lnew = lpp[a];
    const double lnr = (double)(nt - 1) * log(z[a]) + lnew - lp[w];
    const double u = philox_u01(ensemble_counter(step, h, a), seed, kDrawAccept);
    const bool ok = (log(u) < lnr) && isfinite(lnew);   // NaN compares false: rejected
    if (!ok) return;
    double *xa = X + (size_t)nt * (size_t)w;
    const double *p = P + (size_t)nt * (size_t)a;
    for (int64_t k = lane; k < nt; k += 32) xa[k] = p[k];
    if (lane == 0) {
        lp[w] = lnew;
        atomicAdd(n_accept, 1ULL);
    }
}

}  // namespace sfh
