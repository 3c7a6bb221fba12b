// sfh_batched.cuh -- K6: batched-walker log-likelihood.
//
// Per-walker semantics are those of the reference's MCMCModel callable
// (src/fitting/mcmc_sample.jl:12-23): any negative coefficient -> -Inf before any arithmetic
// (:15-19); otherwise composite! (:21) and loglikelihood (:22).  The reference evaluates ONE walker
// per call (W independent gemv's); here W walkers make the composite a dense (Nb x T)·(T x W)
// contraction whose Nb x W result is never materialised: the Poisson term is applied to the FP64
// accumulators in registers and reduced over bins inside the kernel.
//
// FP64 throughout.  tcgen05 has no FP64 kind, so the tensor path for this dtype is mma.sync.m8n8k4.f64 (DMMA): the
// shipped kernel is sfh_batched_logl_mma_kernel below; sfh_batched_logl_kernel (v1, FP64-FMA 8x8 register tiles) is kept
// selectable (SFH_BATCHED_IMPL=fma) as the baseline the DMMA kernel was measured against.  The same file holds K6g, the
// batched gradient GEMM G = M'R used by sfh_eval_fg_batched / sfh_eval_fg_hier_batched.
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
    int64_t nb, nt, W, wld;
    StackLayout lay;
    double eps;
    const double *Xt;    // [nt][wld]
    const double *data;  // [nb]
    double *part;        // [n_bin_tiles][wld]
    double *resid;       // nullable: [nb_padded][wld] residual matrix 1 - n/max(m,eps) for the batched gradient (K6g)
    const struct LogTable *logtab;   // device copy of the log table (fast Poisson epilogue of the DMMA kernel)
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
            ra[u] = (k < p.nt && i < p.nb) ? (double)M[p.lay.off(i, k)] : 0.0;
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

// ------------------------------------------------------------------------------------------
// K6 v2: the same contraction on the FP64 tensor path (mma.sync.m8n8k4.f64 -> DMMA).  ncu on v1 showed the
// kernel compute-bound (FP64 pipe 43 % busy, DRAM ~2 %) with its issue slots and shared-memory wavefronts spent
// on operand delivery (8 LDS.128 + 64 DFMA per thread per k); one DMMA replaces 8 warp-wide DFMA and its
// fragments need 5x fewer shared-memory wavefronts.  tcgen05 has no FP64 kind, so mma.sync is the tensor path
// for this dtype.  Warp tile 32 bins x 8*NBW walkers = 4 x NBW DMMA tiles, BK = 16 templates per slab, 3-stage cp.async
// pipeline; rows padded by 4 doubles => conflict-free fragment loads.
// ------------------------------------------------------------------------------------------
// WN = warps along the walker axis, NBW = 8-walker MMA blocks per warp: CTA tile = 128 bins x 8*NBW*WN walkers, 4*WN warps.
// Shipped shapes: (1, 4) = 32 walkers, 4 warps, 3 CTAs/SM for every W > 16 (never slower than the 64- and 128-wide
// tiles (2, 4) / (4, 4), up to 14 % faster: profiles/r1_experiments.md), (1, 2) / (1, 1) for W <= 16 / W <= 8 -- the
// few-chain batches of sfh_eval_fg_batched, where a wider tile would only multiply zeros.
#ifndef SFH_MMA_STAGES
#define SFH_MMA_STAGES 3
#endif
#ifndef SFH_MMA_BK
#define SFH_MMA_BK 16
#endif
constexpr int kMmaBM = 128, kMmaBK = SFH_MMA_BK, kMmaStages = SFH_MMA_STAGES;

// ---- Poisson epilogue arithmetic -------------------------------------------------------------------------------
// ncu (profiles/r1_experiments.md): with T = 500 the warps of the DMMA kernel spent 42 % of their stall samples in the
// epilogue -- 46 FP64 + ~100 other instructions per (bin, walker) element for one IEEE division and one libdevice log,
// both with slow-path branches that stop the compiler interleaving the 32 elements a thread owns.  The fast path below
// is branch-free: reciprocal by rcp.approx + 2 Newton steps + one correction, and a table-driven log
//     x = 2^k z,  z in [0.6875, 1.375),  c_i = centre of the i-th of 128 sub-intervals (c = 1 for the one starting at 1),
//     log x = k ln2 + log c_i + log1p(r),  r = z / c_i - 1 (one FMA, |r| <= 2^-7), log1p by a degree-8 Taylor polynomial
// (truncation < 2^-56 relative to r; log(1) = 0 exactly, so the `m == n everywhere -> logL = 0 -> -Inf` guard of
// fitting_base.jl:95 is unchanged).  A thread whose inputs leave the fast path's domain (non-finite or out-of-range
// values, negative "counts") takes the exact libdevice path for its whole tile.
constexpr int kLogTabEntries = 128;
struct LogTable { double2 e[kLogTabEntries]; };   // {1/c_i, log c_i}

__device__ __forceinline__ double fast_recip(double m) {   // m in [eps, 1e300)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(m));
    double e = fma(-m, y, 1.0);
    y = fma(y, e, y);
    e = fma(-m, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_div(double n, double m, double y /* ~1/m */) {
    const double q = n * y;
    return fma(fma(-q, m, n), y, q);
}
__device__ __forceinline__ double fast_log(double x, const double2 *__restrict__ tab) {   // x normal, positive, finite
    const long long ix = __double_as_longlong(x);
    const long long tmp = ix - 0x3FE6000000000000LL;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)(tmp >> 52);
    const double z = __longlong_as_double(ix - (tmp & (long long)0xFFF0000000000000ULL));
    const double2 t = tab[i];
    const double r = fma(z, t.x, -1.0);
    const double kd = (double)k;
    double q = fma(r, -1.0 / 8, 1.0 / 7);
    q = fma(r, q, -1.0 / 6);
    q = fma(r, q, 1.0 / 5);
    q = fma(r, q, -1.0 / 4);
    q = fma(r, q, 1.0 / 3);
    q = fma(r, q, -1.0 / 2);
    const double lo = fma(r * r, q, kd * 1.9082149292705877e-10);          // ln2 = hi + lo, hi has 21 trailing zero bits
    return fma(kd, 6.9314718036912382e-01, t.y) + (r + lo);
}
constexpr int kMmaBN = 128, kMmaThreads = 512;  // the (4, 4) shape (host-side grid arithmetic of the walker path)
constexpr int kMmaLdA = kMmaBM + 4;  // doubles; stride = 4 (mod 16) => 16 distinct bank pairs
template <typename S>
__host__ __device__ constexpr size_t mma_smem_bytes(int bn = 128) {
    return (size_t)kMmaStages * kMmaBK * (kMmaLdA * sizeof(S) + (bn + 4) * sizeof(double)) + 4 * bn * sizeof(double) + kLogTabEntries * 16;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <typename S, int WN, int NBW>
__global__ void __launch_bounds__(128 * WN, WN == 1 ? 3 : 1) sfh_batched_logl_mma_kernel(const S *__restrict__ M, const BatchedParams p) {
    constexpr int kWarpN = 8 * NBW, kMmaBN = kWarpN * WN, kMmaThreads = 128 * WN, kMmaLdB = kMmaBN + 4;
    extern __shared__ __align__(16) unsigned char bsm[];
    S *As = reinterpret_cast<S *>(bsm);                                                        // [stages][BK][LdA]
    double *Bs = reinterpret_cast<double *>(bsm + (size_t)kMmaStages * kMmaBK * kMmaLdA * sizeof(S));  // [stages][BK][LdB]
    double *colsum = Bs + (size_t)kMmaStages * kMmaBK * kMmaLdB;                               // [4][BN]
    double2 *ltab = reinterpret_cast<double2 *>(colsum + 4 * kMmaBN);                          // [128] log table
    constexpr int EPV = 16 / sizeof(S);  // A elements per 16-byte cp.async

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < kLogTabEntries; e += kMmaThreads) ltab[e] = p.logtab->e[e];
    const int wm = warp & 3, wn = warp >> 2;  // warp tile origin: bins wm*32, walkers wn*32
    const int64_t n_wt = (p.W + kMmaBN - 1) / kMmaBN;
    const int64_t bt = blockIdx.x / n_wt, wt = blockIdx.x % n_wt;  // walker tiles fastest: the M tile is shared via L2
    const int64_t i0 = bt * kMmaBM, w0 = wt * kMmaBN;
    const int64_t nslab = (p.nt + kMmaBK - 1) / kMmaBK;

    // Loader state is per-thread constant: a thread always copies the same 16-byte column piece (iv / wv) of rows
    // kk0, kk0 + KPI, ... of every slab, and both layouts are linear in the template index (panel: off(i, k) =
    // off(i, 0) + (k << bt_shift)), so a slab costs one multiply-add per cp.async instead of the layout arithmetic
    // (the address chains were the top stall of the 4-warp tile: ncu "wait" 54 k samples vs 24 k math-pipe throttle).
    constexpr int AVR = kMmaBM / EPV;                 // 16-byte pieces per A row
    constexpr int A_KPI = kMmaThreads / AVR;          // A rows covered per pass over the threads
    constexpr int A_NI = kMmaBK / A_KPI;
    static_assert(kMmaThreads % AVR == 0 && kMmaBK % A_KPI == 0 && A_NI >= 1, "A loader shape");
    constexpr int BVR = kMmaBN / 2;                   // 16-byte pieces per B row
    constexpr int B_KPI = (kMmaThreads / BVR) < kMmaBK ? (kMmaThreads / BVR) : kMmaBK;
    constexpr int B_NI = kMmaBK / B_KPI;
    static_assert(kMmaThreads % BVR == 0 && kMmaBK % B_KPI == 0, "B loader shape");
    const int a_iv = (tid % AVR) * EPV, a_kk0 = tid / AVR;
    // rows in [nb, padded) are zero padding; a 16-byte piece never straddles a panel (bt*sizeof(S) >= 64)
    const bool a_iok = (i0 + a_iv) < p.lay.ld;
    const S *a_src = M + (a_iok ? p.lay.off(i0 + a_iv, 0) : 0);
    const int64_t a_kstride = p.lay.panel ? ((int64_t)1 << p.lay.bt_shift) : p.lay.ld;
    S *a_dst = As + (size_t)a_kk0 * kMmaLdA + a_iv;
    const int b_wv = (tid % BVR) * 2, b_kk0 = tid / BVR;
    const bool b_wok = (b_kk0 < kMmaBK) && (w0 + b_wv) < p.wld;   // walkers in [W, wld) hold finite junk: masked later
    const double *b_src = p.Xt + (b_wok ? (w0 + b_wv) : 0);
    double *b_dst = Bs + (size_t)(b_kk0 < kMmaBK ? b_kk0 : 0) * kMmaLdB + b_wv;

    auto issue_slab = [&](int64_t slab, int stage) {
        const int64_t k0 = slab * kMmaBK;
        const bool live = slab < nslab;
#pragma unroll
        for (int n = 0; n < A_NI; ++n) {
            const int64_t k = k0 + a_kk0 + n * A_KPI;
            const bool ok = live && a_iok && (k < p.nt);
            cp_async16(a_dst + ((size_t)stage * kMmaBK + n * A_KPI) * kMmaLdA, a_src + (ok ? k * a_kstride : 0), ok);
        }
        if (b_kk0 < kMmaBK) {
#pragma unroll
            for (int n = 0; n < B_NI; ++n) {
                const int64_t k = k0 + b_kk0 + n * B_KPI;
                const bool ok = live && b_wok && (k < p.nt);
                cp_async16(b_dst + ((size_t)stage * kMmaBK + n * B_KPI) * kMmaLdB, b_src + (ok ? k * p.wld : 0), ok);
            }
        }
        cp_async_commit();
    };

    double acc[4][NBW][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NBW; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int s = 0; s < kMmaStages - 1; ++s) issue_slab(s, s);
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index of this lane
    int stage = 0, fill = kMmaStages - 1;   // ring positions kept as counters (no 64-bit modulo per slab)
    for (int64_t slab = 0; slab < nslab; ++slab) {
        cp_async_wait<kMmaStages - 2>();
        __syncthreads();  // slab `slab` has landed for everyone; the stage refilled below was consumed last iteration
        issue_slab(slab + kMmaStages - 1, fill);
        fill = (fill + 1 == kMmaStages) ? 0 : fill + 1;
        const S *Asl = As + (size_t)stage * kMmaBK * kMmaLdA;
        const double *Bsl = Bs + (size_t)stage * kMmaBK * kMmaLdB;
#pragma unroll
        for (int k4 = 0; k4 < kMmaBK; k4 += 4) {
            double a[4], b[NBW];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = (double)Asl[(k4 + fk) * kMmaLdA + wm * 32 + mb * 8 + fr];
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) b[nb] = Bsl[(k4 + fk) * kMmaLdB + wn * kWarpN + nb * 8 + fr];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) dmma_m8n8k4(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        }
        stage = (stage + 1 == kMmaStages) ? 0 : stage + 1;
    }
    cp_async_wait<0>();

    __syncthreads();   // the log table is in shared memory even when there was no slab to wait for

    // Poisson epilogue: lane holds C[row = mb*8 + lane/4][col = nb*8 + (lane%4)*2 + {0,1}] of its warp tile
    double csum[NBW][2];
#pragma unroll
    for (int nb = 0; nb < NBW; ++nb) csum[nb][0] = csum[nb][1] = 0.0;
    // pass 1: does every element of this thread sit in the fast path's domain?  counts 0 or in [1e-140, 1e140] and
    // composites below 1e140 (NaN fails the comparison) keep n/m and its reciprocal normal numbers.
    double nrow[4];
    bool fast = true;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int64_t i = i0 + wm * 32 + mb * 8 + fr;
        const double n = (i < p.nb) ? p.data[i] : 0.0;
        nrow[mb] = n;
        fast = fast && ((n == 0.0) || (n >= 1e-140 && n <= 1e140));
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) fast = fast && (acc[mb][nb][0] < 1e140) && (acc[mb][nb][1] < 1e140);
    }
    if (fast) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
            const int64_t i = i0 + wm * 32 + mb * 8 + fr;
            const double n = nrow[mb];
            const bool pos = n > 0.0, inb = i < p.nb;
            double rr[NBW][2];
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double m = acc[mb][nb][e];
                    const double mc = (m < p.eps) ? p.eps : m;
                    const double ratio = fast_div(n, mc, fast_recip(mc));
                    const double lq = fast_log(pos ? ratio : 1.0, ltab);
                    const double term = pos ? (n - mc - n * lq) : -mc;  // fitting_base.jl:92
                    csum[nb][e] += inb ? term : 0.0;                    // zero-padding rows beyond nb contribute nothing
                    rr[nb][e] = 1.0 - ratio;                            // fitting_base.jl:279
                }
            if (p.resid && i < p.nb) {  // residual for every (bin, chain): feeds the batched gradient
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    const int64_t w = w0 + wn * kWarpN + nb * 8 + fk * 2;
                    if (w < p.wld) *reinterpret_cast<double2 *>(p.resid + i * p.wld + w) = make_double2(rr[nb][0], rr[nb][1]);
                }
            }
        }
    } else {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
            const int64_t i = i0 + wm * 32 + mb * 8 + fr;
            if (i < p.nb) {
                const double n = p.data[i];
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    csum[nb][0] += poisson_term(acc[mb][nb][0], n, p.eps);
                    csum[nb][1] += poisson_term(acc[mb][nb][1], n, p.eps);
                    if (p.resid) {
                        const int64_t w = w0 + wn * kWarpN + nb * 8 + fk * 2;
                        if (w < p.wld) {
                            const double m0 = acc[mb][nb][0], m1 = acc[mb][nb][1];
                            double2 r;
                            r.x = 1.0 - n / ((m0 < p.eps) ? p.eps : m0);
                            r.y = 1.0 - n / ((m1 < p.eps) ? p.eps : m1);
                            *reinterpret_cast<double2 *>(p.resid + i * p.wld + w) = r;
                        }
                    }
                }
            }
        }
    }
    // sum over the 8 lanes that share lane%4 (the rows of the fragment): fixed xor tree
#pragma unroll
    for (int nb = 0; nb < NBW; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = csum[nb][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            csum[nb][e] = v;
        }
    if (fr == 0) {
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) colsum[wm * kMmaBN + wn * kWarpN + nb * 8 + fk * 2 + e] = csum[nb][e];
    }
    __syncthreads();
    if (tid < kMmaBN) {
        const int64_t w = w0 + tid;
        if (w < p.W)
            p.part[bt * p.wld + w] = (colsum[tid] + colsum[kMmaBN + tid]) + (colsum[2 * kMmaBN + tid] + colsum[3 * kMmaBN + tid]);
    }
}

// ------------------------------------------------------------------------------------------
// K6g: batched gradient  G[T x C] = M' (T x Nb) * R (Nb x C)  for C coefficient vectors at once (multi-chain HMC:
// the reference runs chains on separate threads, hmc_sample.jl:123-141 / generic_fitting.jl:617-626; here one device
// pass serves them all).  Split-K over bins: CTA (tt, sp) owns 128 templates x all C chains x a contiguous range of
// bin slabs and writes a partial; sfh_bgrad_reduce_kernel sums the partials in fixed order.  DMMA m8n8k4, 4 warps,
// warp tile 32 templates x C (C <= 64), slabs of 16 bins through a 3-stage cp.async pipeline.
// ------------------------------------------------------------------------------------------
constexpr int kBgBM = 128, kBgBK = 16, kBgStages = 3, kBgThreads = 128, kBgMaxC = 64;
constexpr int kBgLdA = kBgBK + 4;  // doubles/floats per template row of a slab: stride = 4 (mod 16) => conflict-free
struct BGradParams {
    int64_t nb, nt, wld;   // wld: row length of the residual matrix (>= C, multiple of 8)
    int32_t C, nsplit;
    StackLayout lay;
    const double *resid;   // [nb_padded][wld]
    double *gpart;         // [nsplit][nt][wld]
};
template <typename S>
__host__ __device__ constexpr size_t bgrad_smem_bytes(int C) {
    return (size_t)kBgStages * (kBgBM * kBgLdA * sizeof(S) + kBgBK * (size_t)(C + 4) * sizeof(double));
}

template <typename S, int NB /* n-blocks of 8 chains */>
__global__ void __launch_bounds__(kBgThreads) sfh_bgrad_mma_kernel(const S *__restrict__ M, const BGradParams p) {
    extern __shared__ __align__(16) unsigned char bsm[];
    constexpr int C = NB * 8, LdB = C + 4;
    constexpr int EPV = 16 / sizeof(S);
    S *As = reinterpret_cast<S *>(bsm);                                                      // [stages][BM][LdA]
    double *Bs = reinterpret_cast<double *>(bsm + (size_t)kBgStages * kBgBM * kBgLdA * sizeof(S));  // [stages][BK][LdB]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;
    const int64_t j0 = (int64_t)blockIdx.x * kBgBM;
    const int64_t nslab_all = (p.nb + kBgBK - 1) / kBgBK;
    const int64_t per = (nslab_all + p.nsplit - 1) / p.nsplit;
    const int64_t s0 = (int64_t)blockIdx.y * per, s1 = (s0 + per < nslab_all) ? s0 + per : nslab_all;

    // per-thread constant loader state (cf. the logL kernel): a thread copies the same 16-byte piece (bins kv..) of
    // templates m0, m0 + A_MPI, ... and the same column pair of rows k0, k0 + B_KPI, ... of every slab
    constexpr int A_VEC = kBgBK / EPV;                 // 16-byte pieces per template row of a slab
    constexpr int A_MPI = kBgThreads / A_VEC;          // templates covered per pass over the threads
    constexpr int A_NI = kBgBM / A_MPI;
    static_assert(kBgThreads % A_VEC == 0 && kBgBM % A_MPI == 0, "A loader shape");
    constexpr int B_VEC = C / 2;
    constexpr int B_KPI = (kBgThreads / B_VEC) < kBgBK ? (kBgThreads / B_VEC) : kBgBK;
    constexpr int B_NI = kBgBK / B_KPI;
    static_assert(kBgThreads % B_VEC == 0 && kBgBK % B_KPI == 0, "B loader shape");
    const int a_kv = (tid % A_VEC) * EPV, a_m0 = tid / A_VEC;
    const int64_t a_sj = p.lay.panel ? ((int64_t)1 << p.lay.bt_shift) : p.lay.ld;   // stride between templates
    S *a_dst = As + (size_t)a_m0 * kBgLdA + a_kv;
    const int b_cv = (tid % B_VEC) * 2, b_k0 = tid / B_VEC;
    const bool b_on = (b_k0 < kBgBK) && (b_cv < p.wld);   // the kernel's C (multiple of 8, <= 64) may exceed the row length
    double *b_dst = Bs + (size_t)(b_k0 < kBgBK ? b_k0 : 0) * LdB + b_cv;

    auto issue = [&](int64_t slab, int stage) {
        const int64_t i0 = slab * kBgBK;
        const bool live = slab < s1;
        // A: 128 templates x 16 bins; 16-byte pieces along bins (contiguous inside a panel row)
        const int64_t ia = i0 + a_kv;
        const bool a_ok = live && ia < p.lay.ld;
        const S *a_src = M + (a_ok ? p.lay.off(ia, j0 + a_m0) : 0);
#pragma unroll
        for (int n = 0; n < A_NI; ++n) {
            const bool ok = a_ok && (j0 + a_m0 + n * A_MPI) < p.nt;
            cp_async16(a_dst + ((size_t)stage * kBgBM + n * A_MPI) * kBgLdA, a_src + (ok ? (int64_t)(n * A_MPI) * a_sj : 0), ok);
        }
        // B: 16 bins x C chains of the residual matrix (rows beyond nb were never written: masked by `i < nb`)
        if (b_k0 < kBgBK) {
#pragma unroll
            for (int n = 0; n < B_NI; ++n) {
                const int64_t i = i0 + b_k0 + n * B_KPI;
                const bool ok = live && b_on && i < p.nb;
                cp_async16(b_dst + ((size_t)stage * kBgBK + n * B_KPI) * LdB, p.resid + (ok ? i * p.wld + b_cv : 0), ok);
            }
        }
        cp_async_commit();
    };

    double acc[4][NB][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int s = 0; s < kBgStages - 1; ++s) issue(s0 + s, s);
    int stage = 0, fill = kBgStages - 1;
    for (int64_t slab = s0; slab < s1; ++slab) {
        cp_async_wait<kBgStages - 2>();
        __syncthreads();
        issue(slab + kBgStages - 1, fill);
        fill = (fill + 1 == kBgStages) ? 0 : fill + 1;
        const S *Asl = As + (size_t)stage * kBgBM * kBgLdA;
        const double *Bsl = Bs + (size_t)stage * kBgBK * LdB;
#pragma unroll
        for (int k4 = 0; k4 < kBgBK; k4 += 4) {
            double a[4], b[NB];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = (double)Asl[(warp * 32 + mb * 8 + fr) * kBgLdA + k4 + fk];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) b[nb] = Bsl[(k4 + fk) * LdB + nb * 8 + fr];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) dmma_m8n8k4(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        }
        stage = (stage + 1 == kBgStages) ? 0 : stage + 1;
    }
    cp_async_wait<0>();
    // C fragment: row (template) = mb*8 + lane/4, cols (chains) = nb*8 + (lane%4)*2 + {0,1}
    double *out = p.gpart + (size_t)blockIdx.y * p.nt * p.wld;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int64_t j = j0 + warp * 32 + mb * 8 + fr;
        if (j < p.nt) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                if (nb * 8 + fk * 2 < p.wld) {
                    double2 v;
                    v.x = acc[mb][nb][0];
                    v.y = acc[mb][nb][1];
                    *reinterpret_cast<double2 *>(out + j * p.wld + nb * 8 + fk * 2) = v;
                }
            }
        }
    }
}

// G[j + nt*c] (column-major T x C, like the coefficient matrix) = sum over splits, fixed order
__global__ void sfh_bgrad_reduce_kernel(const double *__restrict__ gpart, int nsplit, int64_t nt, int64_t wld, int64_t C,
                                        double *__restrict__ G, int64_t ostride, int64_t ooff) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt * C) return;
    const int64_t j = e % nt, c = e / nt;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += gpart[((size_t)sp * nt + j) * wld + c];
    G[c * ostride + ooff + j] = s;   // ostride = nt, ooff = 0: compact T x C; (1 + nt, 1): the hierarchical epilogue's [logL, G] rows
}

// logL[w] = sum over bin tiles; raw sums (guards applied after any all-reduce).  256 threads = 8 walkers x 32 slices:
// slice q adds tiles q, q+32, ... in order, then a fixed shared-memory tree joins the slices (deterministic, and not a
// 469-long dependent chain per walker).
__global__ void __launch_bounds__(256) sfh_batched_reduce_kernel(const double *__restrict__ part, int64_t n_bt, int64_t W, int64_t wld,
                                                                  double *__restrict__ out) {
    __shared__ double sh[32][8];
    const int wl = threadIdx.x & 7, q = threadIdx.x >> 3;
    const int64_t w = (int64_t)blockIdx.x * 8 + wl;
    double s = 0.0;
    if (w < W)
        for (int64_t b = q; b < n_bt; b += 32) s += part[b * wld + w];
    sh[q][wl] = s;
    __syncthreads();
    for (int h = 16; h >= 1; h >>= 1) {
        if (q < h) sh[q][wl] += sh[q + h][wl];
        __syncthreads();
    }
    if (q == 0 && w < W) out[w] = sh[0][wl];
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
