// sfh_batched.cuh -- K6: batched-walker log-likelihood.
//
// Per-walker semantics are those of the reference's MCMCModel callable
// (src/fitting/mcmc_sample.jl:12-23): any negative coefficient -> -Inf before any arithmetic
// (:15-19); otherwise composite! (:21) and loglikelihood (:22).  The reference evaluates ONE walker
// per call (W independent gemv's); here W walkers make the composite a dense (Nb x T)·(T x W)
// contraction whose Nb x W result is never materialised: the Poisson term is applied to the FP64
// accumulators in registers and reduced over bins inside the kernel.
//
// FP64 throughout (tcgen05 has no FP64 kind; B200's FP64 tensor rate equals its FP64 FMA rate, so
// the contraction runs on the FP64 pipe with an 8x8 register tile per thread).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sfh_small.cuh"

namespace sfh {

constexpr int kBwBM = 128;  // bins per CTA tile
constexpr int kBwBN = 128;  // walkers per CTA tile
constexpr int kBwBK = 8;    // templates per slab (2 x 2 x 8 KB of static shared memory)
constexpr int kBwThreads = 256;

// Xt[k][w] = X[k + T*w]  (walker-contiguous copy for coalesced slab loads) and
// neg[w] = any(X[:,w] < 0)   (mcmc_sample.jl:15-19).  One warp per walker.
__global__ void sfh_walker_prep_kernel(const double *__restrict__ X, int64_t nt, int64_t W, int64_t wld,
                                       double *__restrict__ Xt, int32_t *__restrict__ neg) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= W) return;
    bool any = false;
    for (int64_t k = lane; k < nt; k += 32) {
        const double v = X[k + nt * w];
        any |= (v < 0.0);
        Xt[k * wld + w] = v;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) neg[w] = any ? 1 : 0;
}

struct BatchedParams {
    int64_t nb, nt, W, ld, wld;
    double eps;
    const double *Xt;    // [nt][wld]
    const double *data;  // [nb]
    double *part;        // [n_bin_tiles][wld]
};

template <typename S>
__global__ void __launch_bounds__(kBwThreads) sfh_batched_logl_kernel(const S *__restrict__ M, const BatchedParams p) {
    __shared__ __align__(16) double As[2][kBwBK][kBwBM];
    __shared__ __align__(16) double Bs[2][kBwBK][kBwBN];
    const int tid = threadIdx.x;
    const int tx = tid & 15;  // bins:    rows tx*4..+4 and 64+tx*4..+4
    const int ty = tid >> 4;  // walkers: cols ty*4..+4 and 64+ty*4..+4
    const int64_t n_wt = (p.W + kBwBN - 1) / kBwBN;
    const int64_t bt = blockIdx.x / n_wt, wt = blockIdx.x % n_wt;  // walker tiles fastest: M tile shared via L2
    const int64_t i0 = bt * kBwBM, w0 = wt * kBwBN;

    // slab loaders: A: 8 templates x 128 bins = 1024 values -> 4 per thread;  B: same
    const int a_i = tid & 127, a_k = tid >> 7;  // rows a_k + 2*u
    double ra[4], rb[4];
    auto load_slab = [&](int64_t k0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t k = k0 + a_k + 2 * u;
            const int64_t i = i0 + a_i;
            ra[u] = (k < p.nt && i < p.nb) ? (double)M[i + k * p.ld] : 0.0;
            const int64_t w = w0 + a_i;
            rb[u] = (k < p.nt && w < p.W) ? p.Xt[k * p.wld + w] : 0.0;
        }
    };
    auto store_slab = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            As[buf][a_k + 2 * u][a_i] = ra[u];
            Bs[buf][a_k + 2 * u][a_i] = rb[u];
        }
    };

    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

    const int64_t nslab = (p.nt + kBwBK - 1) / kBwBK;
    load_slab(0);
    store_slab(0);
    __syncthreads();
    for (int64_t s = 0; s < nslab; ++s) {
        const int buf = (int)(s & 1);
        if (s + 1 < nslab) load_slab((s + 1) * kBwBK);
#pragma unroll
        for (int k = 0; k < kBwBK; ++k) {
            double a[8], b[8];
            const double2 a0 = *reinterpret_cast<const double2 *>(&As[buf][k][tx * 4]);
            const double2 a1 = *reinterpret_cast<const double2 *>(&As[buf][k][tx * 4 + 2]);
            const double2 a2 = *reinterpret_cast<const double2 *>(&As[buf][k][64 + tx * 4]);
            const double2 a3 = *reinterpret_cast<const double2 *>(&As[buf][k][64 + tx * 4 + 2]);
            const double2 b0 = *reinterpret_cast<const double2 *>(&Bs[buf][k][ty * 4]);
            const double2 b1 = *reinterpret_cast<const double2 *>(&Bs[buf][k][ty * 4 + 2]);
            const double2 b2 = *reinterpret_cast<const double2 *>(&Bs[buf][k][64 + ty * 4]);
            const double2 b3 = *reinterpret_cast<const double2 *>(&Bs[buf][k][64 + ty * 4 + 2]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y; a[6] = a3.x; a[7] = a3.y;
            b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y; b[4] = b2.x; b[5] = b2.y; b[6] = b3.x; b[7] = b3.y;
#pragma unroll
            for (int x = 0; x < 8; ++x)
#pragma unroll
                for (int y = 0; y < 8; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        if (s + 1 < nslab) {
            store_slab(buf ^ 1);
            __syncthreads();
        }
    }

    // Poisson epilogue on the register tile; reduce over this thread's 8 bins
    double wsum[8];
#pragma unroll
    for (int y = 0; y < 8; ++y) wsum[y] = 0.0;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        const int64_t i = i0 + ((x < 4) ? tx * 4 + x : 64 + tx * 4 + (x - 4));
        if (i < p.nb) {
            const double n = p.data[i];
#pragma unroll
            for (int y = 0; y < 8; ++y) wsum[y] += poisson_term(acc[x][y], n, p.eps);
        }
    }
    // reduce over the 16 tx lanes that share ty (fixed xor tree inside the half-warp)
#pragma unroll
    for (int y = 0; y < 8; ++y) {
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) wsum[y] += __shfl_xor_sync(0xffffffffu, wsum[y], off);
    }
    if (tx == 0) {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int64_t w = w0 + ((y < 4) ? ty * 4 + y : 64 + ty * 4 + (y - 4));
            if (w < p.W) p.part[bt * p.wld + w] = wsum[y];
        }
    }
}

// logL[w] = sum over bin tiles (fixed order); raw sums (guards applied after any all-reduce)
__global__ void sfh_batched_reduce_kernel(const double *__restrict__ part, int64_t n_bt, int64_t W, int64_t wld,
                                          double *__restrict__ out) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double s = 0.0;
    for (int64_t b = 0; b < n_bt; ++b) s += part[b * wld + w];
    out[w] = s;
}

// mcmc_sample.jl:15-19 (negative -> -Inf) and fitting_base.jl:95 (== 0 -> -Inf)
__global__ void sfh_batched_guard_kernel(double *__restrict__ logl, const int32_t *__restrict__ neg, int64_t W) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const double v = logl[w];
    const double ninf = __longlong_as_double(0xfff0000000000000LL);
    logl[w] = neg[w] ? ninf : ((v != 0.0) ? v : ninf);
}

}  // namespace sfh
