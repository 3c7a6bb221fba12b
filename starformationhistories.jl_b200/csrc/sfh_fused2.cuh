// sfh_fused2.cuh -- K4 v2: the fused composite -> residual -> transposed-gradient kernel as a WARP-SPECIALISED
// STREAM (sm_100a).  Same arithmetic, same one-pass-over-HBM contract as sfh_fused.cuh:
//     composite!            src/fitting/fitting_base.jl:55-65   (gemv 'N')
//     grad-loglikelihood!   src/fitting/fitting_base.jl:265-285 (residual + gemv 'T')
//     loglikelihood         src/fitting/fitting_base.jl:84-96   (per-bin Poisson term, now inside this kernel)
// as sequenced by fg! (src/fitting/solvers.jl:20-38).
//
// What round 1 measured (profiles/r1_experiments.md, r2_experiments.md): in sfh_fused.cuh the SAME warps run
// pass A -> cluster exchange -> pass B per tile, the tile stays resident in the ring meanwhile, and the ring
// holds barely more than one tile -- so the HBM stream drains at every tile boundary (config 5: 0.73-0.79 of
// the roofline) and "two tiles in flight in the same warps" (sfh_fused_pipe.cuh, deleted) was slower still.
// Here the three phases are three ROLES that never wait for each other except through mbarriers:
//   producer (1 warp, 1 lane)   1-D bulk copies (cp.async.bulk, 16 KB stages) into a ring of NS stages; the
//                               device layout is bin-major panels of BT bins, so a CTA's slice of a tile is
//                               ONE contiguous byte range -- no tensor map at all.
//   A warps  (8)                coefficients live in REGISTERS (<= 20 per lane); each lane FMAs its 16 bytes of
//                               every chunk into VEC composite partials; per tile one warp reduction and one
//                               mbarrier arrive.  They never touch the ring's empty barriers (WANT_G).
//   reducer  (1 warp)           sums the 8 warp partials, exchanges CTA partials through DSMEM (st.async +
//                               complete_tx) when the cluster has more than one CTA, forms the residual
//                               r = 1 - n/max(m,eps) (fitting_base.jl:277-279), publishes it, and -- off the
//                               critical path -- writes the composite and accumulates the Poisson term.
//   B warps  (8)                gradient partials in registers across ALL tiles of the CTA; re-read the same
//                               bytes from SHARED memory once r is published, then release the stage.
// A runs ahead of B by as much as the ring allows (tiles are small: BT = 16 bytes' worth of bins per lane row
// times LPR lanes), so the exchange latency is hidden by streaming, not by a second resident CTA.
// Deterministic: fixed-order sums everywhere, one plain store per (cluster, template), no atomics.
#pragma once
#include "sfh_ptx.cuh"

namespace sfh {

constexpr int kV2A = 8;                                   // A (composite) warps
constexpr int kV2B = 8;                                   // B (gradient) warps
constexpr int kV2Threads = (kV2A + kV2B + 2) * 32;        // + producer warp + reducer warp
constexpr int kV2KMax = 20;                               // chunks per CTA per tile (per-lane register arrays)
constexpr int kV2G = 4;                                   // chunks per pipeline stage (one bulk copy)
constexpr uint32_t kV2Chunk = kV2A * 32 * 16;             // one 16-byte vector per A lane = 4 KB
constexpr uint32_t kV2Stage = kV2G * kV2Chunk;            // 16 KB
constexpr int kV2SMax = kV2KMax / kV2G;                   // stages per tile at most
constexpr int kV2DS = 8;                                  // exchange slots (tiles between A and B); ring <= (DS-1) tiles

struct Fused2Params {
    int64_t nb;           // bins in this shard
    int64_t nt;           // templates
    int32_t kt;           // chunks per CTA per tile (<= kV2KMax)
    int32_t ns;           // ring stage slots (<= (kV2DS-1) * stages per tile: bounds how far A can run ahead of B)
    int32_t n_tiles;      // ceil(nb / BT)
    int32_t evict_first;  // L2 evict_first policy on the stack stream
    int32_t pdl_early;    // 1: release the dependent (finalize) launch at kernel start
    int32_t keep_stages;  // the first keep_stages stages of every CTA's tile sequence are loaded with the L2 evict_last policy
    double eps;           // clamp (fitting_base.jl:90,277)
    const void *M;        // device stack, bin-major panels [n_tiles][nt][BT]
    const double *coeffs; // [nt]
    const double *data;   // [nb]
    double *composite;    // [nb] out: M*coeffs (unclamped)
    double *residual;     // [nb] out (nullable)
    double *gpart;        // [n_clusters][gstride] out
    double *lpart;        // [n_clusters] out: per-cluster sum of the Poisson terms (raw logL partial)
    int64_t gstride;
    unsigned long long *epoch_ptr;   // nullable: the exchange epoch of the finalize kernel that follows (FinalizeParams::epoch_ptr), bumped
                                     // here -- by ONE thread, after this grid's griddepcontrol.wait, i.e. after the previous finalize
                                     // kernel has completed -- so that the sharded finalize kernel needs no last block for it
};

struct Fused2Smem {
    uint32_t ring_off, red_off, xbuf_off, rbuf_off, bar_off, total;
    __host__ __device__ static Fused2Smem make(int ns, int bt, int cluster) {
        Fused2Smem s;
        s.ring_off = 0;
        s.red_off = (uint32_t)ns * kV2Stage;
        s.xbuf_off = s.red_off + kV2DS * kV2A * bt * 8;
        s.rbuf_off = s.xbuf_off + kV2DS * cluster * bt * 8;
        s.bar_off = s.rbuf_off + kV2DS * bt * 8;
        s.total = s.bar_off + (2 * ns + 3 * kV2DS) * 8;
        return s;
    }
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d_hint(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}

// FAST (Float32 stacks whose elements are all finite with the sign bit clear -- checked once at creation): the float's bits
// are dropped into a double WITHOUT re-biasing the exponent: one 32 x 32 -> 64 multiply by 2^29 puts the 23 mantissa bits and
// the 8 exponent bits where a double keeps them, giving exactly x * 2^-896 (also for float denormals, which land on double
// denormals, and for 0).  The factor 2^896 is folded into the coefficients (A warps) and the residual (B warps), so every
// product m*c and m*r is bit-identical to the converted one.  Why: cvt.f64.f32 runs on the XU pipe at 16 lanes/clk/SM, and at
// two conversions per element ncu shows that pipe 78 % busy at 1.9 GHz -- under a sustained run's power cap (1.6 GHz) the
// conversions, not HBM, bound the Float32 kernel (profiles/r2_experiments.md).
constexpr double kV2FastScale = 0x1p896;
template <typename S, bool FAST>
__device__ __forceinline__ void unpack2(const vec16 &v, double (&out)[16 / sizeof(S)]);
template <>
__device__ __forceinline__ void unpack2<double, false>(const vec16 &v, double (&out)[2]) {
    out[0] = __hiloint2double(v.y, v.x);
    out[1] = __hiloint2double(v.w, v.z);
}
template <>
__device__ __forceinline__ void unpack2<float, false>(const vec16 &v, double (&out)[4]) {
    out[0] = (double)__uint_as_float(v.x);
    out[1] = (double)__uint_as_float(v.y);
    out[2] = (double)__uint_as_float(v.z);
    out[3] = (double)__uint_as_float(v.w);
}
template <>
__device__ __forceinline__ void unpack2<float, true>(const vec16 &v, double (&out)[4]) {
    out[0] = __longlong_as_double((long long)((unsigned long long)v.x * 536870912ull));
    out[1] = __longlong_as_double((long long)((unsigned long long)v.y * 536870912ull));
    out[2] = __longlong_as_double((long long)((unsigned long long)v.z * 536870912ull));
    out[3] = __longlong_as_double((long long)((unsigned long long)v.w * 536870912ull));
}

// Poisson log-likelihood-ratio term, fitting_base.jl:90-92 (`ifelse` = select; NaN propagates like Julia's scalar max)
__device__ __forceinline__ double poisson_term2(double m, double n, double eps) {
    const double mc = (m < eps) ? eps : m;
    return (n > 0.0) ? (n - mc - n * log(n / mc)) : -mc;
}

// LPR = lanes per template row of a chunk: BT = LPR * (16 / sizeof(S)) bins per tile; RPC = 256 / LPR templates per chunk.
template <typename S, int LPR, bool WANT_G, bool FAST>
__global__ void __launch_bounds__(kV2Threads, 1) sfh_fg_fused2_kernel(const Fused2Params p) {
    static_assert(!FAST || sizeof(S) == 4, "FAST is the Float32 conversion-free unpack");
    constexpr int VEC = 16 / sizeof(S), BT = VEC * LPR, RPW = 32 / LPR, RPC = RPW * kV2A;
    constexpr int NBL = (BT + 31) / 32;   // bins per reducer lane
    static_assert(LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "LPR must be a power of two");
    static_assert(RPC * BT * sizeof(S) == kV2Chunk, "chunk = one 16-byte vector per A lane");

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t q = cluster_ctarank();
    const uint32_t C = cluster_nctarank();
    const uint32_t cl = cluster_id_x();
    const uint32_t ncl = cluster_nid_x();
    const int kt = p.kt, NS = p.ns;
    const int nst = (kt + kV2G - 1) / kV2G;
    const Fused2Smem L = Fused2Smem::make(NS, BT, (int)C);

    double *red = reinterpret_cast<double *>(smem + L.red_off);    // [DS][A warps][BT]
    double *xbuf = reinterpret_cast<double *>(smem + L.xbuf_off);  // [DS][C][BT]
    double *rbuf = reinterpret_cast<double *>(smem + L.rbuf_off);  // [DS][BT]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint64_t *empty = full + NS;
    uint64_t *redbar = empty + NS;     // [DS] A warps -> reducer
    uint64_t *xbar = redbar + kV2DS;   // [DS] cluster exchange (tx bytes)
    uint64_t *rbar = xbar + kV2DS;     // [DS] reducer -> B warps

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], WANT_G ? kV2B : kV2A);
        }
        for (int i = 0; i < kV2DS; ++i) {
            mbar_init(&redbar[i], kV2A);
            mbar_init(&xbar[i], 1);
            mbar_init(&rbar[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // every CTA's barriers must be initialised before any peer signals them
    cluster_arrive();
    cluster_wait();
    // PDL.  Everything above overlaps the previous kernel's tail, and so does the PRODUCER: the stack is immutable, so it starts
    // filling the ring while the predecessor (coefficient upload kernel / hierarchical prologue / previous finalize) is still
    // running -- those release this launch at their start.  Every other role waits: the A warps read the predecessor's
    // coefficients, the reducer and the B warps overwrite buffers the previous finalize reads.
    // (pdl_early: release the finalize launch at once as well; measured: its waiting blocks occupy SMs the next evaluation's
    // CTAs want, slower in a sustained loop -- off by default, profiles/r2_experiments.md section 6.)
    if (p.pdl_early) griddep_launch_dependents();
    if (warp != kV2A + kV2B) griddep_wait();

    const int my_tiles = (p.n_tiles > (int)cl) ? (p.n_tiles - (int)cl + (int)ncl - 1) / (int)ncl : 0;
    const uint32_t ring_base = smem_u32(smem + L.ring_off);

    if (warp == kV2A + kV2B) {
        // ================= producer: one bulk copy per stage =================
        if (lane == 0) {
            // L2 residency between evaluations: a stack larger than L2 is streamed evict_first, EXCEPT the head of every CTA's
            // tile sequence (keep_stages stages, ~half of L2 in total), which is loaded evict_last and therefore still there
            // when the next evaluation of the same stack starts: those bytes never touch HBM again, and the ring's first fill
            // comes at L2 latency.  (The stack is immutable and a fit evaluates it thousands of times.)
            const uint64_t pol = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
            int stage_no = 0;
            constexpr int64_t row_bytes = (int64_t)BT * sizeof(S);
            const int64_t row0 = (int64_t)q * kt * RPC;                      // first template of this CTA's slice
            int64_t rows_mine = p.nt - row0;
            rows_mine = rows_mine < 0 ? 0 : (rows_mine > (int64_t)kt * RPC ? (int64_t)kt * RPC : rows_mine);
            const char *M = reinterpret_cast<const char *>(p.M);
            int slot = 0;
            uint32_t round = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int64_t tile = (int64_t)cl + (int64_t)it * ncl;
                const char *base = M + (tile * p.nt + row0) * row_bytes;
                for (int s = 0; s < nst; ++s) {
                    int64_t rows = rows_mine - (int64_t)s * kV2G * RPC;
                    rows = rows < 0 ? 0 : (rows > kV2G * RPC ? kV2G * RPC : rows);
                    if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1u);
                    if (rows > 0) {
                        const uint32_t bytes = (uint32_t)(rows * row_bytes);
                        mbar_arrive_expect_tx(&full[slot], bytes);
                        const uint32_t dst = ring_base + (uint32_t)slot * kV2Stage;
                        const char *src = base + (int64_t)s * kV2G * RPC * row_bytes;
                        if (p.evict_first) bulk_load_1d_hint(dst, src, bytes, smem_u32(&full[slot]), stage_no < p.keep_stages ? pol_keep : pol);
                        else bulk_load_1d(dst, src, bytes, smem_u32(&full[slot]));
                    } else {
                        mbar_arrive(&full[slot]);   // a stage wholly past the last template: nothing to fetch
                    }
                    ++stage_no;
                    if (++slot == NS) { slot = 0; ++round; }
                }
            }
        }
    } else if (warp < kV2A + kV2B) {
        // ================= A and B warps share the lane -> (template, bins) mapping =================
        const int w = (warp < kV2A) ? warp : warp - kV2A;
        const int bl = lane % LPR;   // which VEC-bin group of the tile this lane owns
        const int rw = lane / LPR;   // which template row of the warp's RPW rows
        const uint32_t lane_off = (uint32_t)(w * 32 + lane) * 16u;
        // the lane's templates: j(k) = q*kt*RPC + k*RPC + w*RPW + rw   (fixed for the whole kernel)
        const int64_t j0 = (int64_t)q * kt * RPC + w * RPW + rw;
        // chunks whose template exists for this lane (stale shared memory past the last template is never used)
        int kv = 0;
        if (j0 < p.nt) {
            const int64_t n = (p.nt - j0 + RPC - 1) / RPC;
            kv = n > kt ? kt : (int)n;
        }
        int slot = 0;
        uint32_t phase = 0;

        if (warp < kV2A) {
            // ---------------- pass A: composite partials ----------------
            double creg[kV2KMax];
#pragma unroll
            for (int k = 0; k < kV2KMax; ++k) creg[k] = (k < kv) ? __ldg(p.coeffs + j0 + (int64_t)k * RPC) * (FAST ? kV2FastScale : 1.0) : 0.0;
            for (int it = 0; it < my_tiles; ++it) {
                const int xs = it & (kV2DS - 1);
                double acc[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
#pragma unroll
                for (int s = 0; s < kV2SMax; ++s) {
                    if (s < nst) {
                        mbar_wait(&full[slot], phase);
                        const uint32_t sbase = ring_base + (uint32_t)slot * kV2Stage + lane_off;
                        vec16 v[kV2G];
#pragma unroll
                        for (int u = 0; u < kV2G; ++u)
                            if (s * kV2G + u < kv) v[u] = lds128(sbase + (uint32_t)u * kV2Chunk);
#pragma unroll
                        for (int u = 0; u < kV2G; ++u) {
                            if (s * kV2G + u < kv) {
                                double m[VEC];
                                unpack2<S, FAST>(v[u], m);
#pragma unroll
                                for (int e = 0; e < VEC; ++e) acc[e] = fma(m[e], creg[s * kV2G + u], acc[e]);
                            }
                        }
                        if (!WANT_G) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&empty[slot]);
                        }
                        if (++slot == NS) { slot = 0; phase ^= 1u; }
                    }
                }
                // lanes that own the same bins (different template rows) combine
#pragma unroll
                for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
                }
                // red[xs] was last used by tile it - DS: the reducer must have consumed it (it publishes rbar after reading red).
                // With the gradient pass the ring bound (NS <= (DS-1) tiles) already implies this; without it nothing else does.
                if (it >= kV2DS) mbar_wait(&rbar[xs], (uint32_t)(it / kV2DS - 1) & 1u);
                if (rw == 0) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) red[(xs * kV2A + w) * BT + bl * VEC + e] = acc[e];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&redbar[xs]);
            }
        } else if (WANT_G) {
            // ---------------- pass B: gradient partials from the same bytes ----------------
            double gacc[kV2KMax];
#pragma unroll
            for (int k = 0; k < kV2KMax; ++k) gacc[k] = 0.0;
            for (int it = 0; it < my_tiles; ++it) {
                const int xs = it & (kV2DS - 1);
                mbar_wait(&rbar[xs], (uint32_t)(it / kV2DS) & 1u);
                double r[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) r[e] = rbuf[xs * BT + bl * VEC + e] * (FAST ? kV2FastScale : 1.0);
#pragma unroll
                for (int s = 0; s < kV2SMax; ++s) {
                    if (s < nst) {
                        // already complete (the A warps consumed it); observed here so that the bulk copy's writes are
                        // ordered before this warp's reads as well
                        mbar_wait(&full[slot], phase);
                        const uint32_t sbase = ring_base + (uint32_t)slot * kV2Stage + lane_off;
                        vec16 v[kV2G];
#pragma unroll
                        for (int u = 0; u < kV2G; ++u)
                            if (s * kV2G + u < kv) v[u] = lds128(sbase + (uint32_t)u * kV2Chunk);
#pragma unroll
                        for (int u = 0; u < kV2G; ++u) {
                            if (s * kV2G + u < kv) {
                                double m[VEC];
                                unpack2<S, FAST>(v[u], m);
                                double g = gacc[s * kV2G + u];
#pragma unroll
                                for (int e = 0; e < VEC; ++e) g = fma(m[e], r[e], g);
                                gacc[s * kV2G + u] = g;
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[slot]);
                        if (++slot == NS) { slot = 0; phase ^= 1u; }
                    }
                }
            }
            // one store per (cluster, template)
#pragma unroll
            for (int k = 0; k < kV2KMax; ++k) {
                if (k < kt) {
                    double g = gacc[k];
#pragma unroll
                    for (int off = 1; off < LPR; off <<= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
                    const int64_t j = j0 + (int64_t)k * RPC;
                    if (bl == 0 && j < p.nt) p.gpart[(int64_t)cl * p.gstride + j] = g;
                }
            }
        }
    } else {
        // ================= reducer: partials -> composite -> residual; Poisson term off the critical path =================
        // The reducer is ONE warp per SM and a tile can be as short as 0.8 us, so its per-tile dependency chain is kept minimal:
        // the observed counts are fetched one tile ahead, and when a tile has fewer bins than the warp has lanes the Poisson terms
        // (a ~400-cycle FP64 log each) are queued -- (m, n) pairs parked in idle lanes by shuffle -- and evaluated 32 at a time.
        // (Round 2, first version: one log per tile on 2 of 32 lanes kept up at 1.9 GHz but not under a sustained run's power
        // cap: config 3 with 2-bin tiles went from 181 us isolated to 202 us per step back to back.)
        constexpr int QT = (BT < 32) ? 32 / BT : 1;   // tiles per queue flush
        double lacc = 0.0, qm = 0.0, qn = 0.0;
        bool qvalid = false;
        int qcount = 0;
        double n_next[NBL];
#pragma unroll
        for (int i = 0; i < NBL; ++i) {
            const int b = lane + 32 * i;
            const int64_t bin = (int64_t)cl * BT + b;
            n_next[i] = (my_tiles > 0 && b < BT && bin < p.nb) ? __ldg(p.data + bin) : 0.0;
        }
        for (int it = 0; it < my_tiles; ++it) {
            const int64_t tile = (int64_t)cl + (int64_t)it * ncl;
            const int xs = it & (kV2DS - 1);
            const uint32_t par = (uint32_t)(it / kV2DS) & 1u;
            double n_obs[NBL], m[NBL];
#pragma unroll
            for (int i = 0; i < NBL; ++i) {
                n_obs[i] = n_next[i];
                const int b = lane + 32 * i;
                const int64_t bin = (tile + ncl) * BT + b;
                n_next[i] = (it + 1 < my_tiles && b < BT && bin < p.nb) ? __ldg(p.data + bin) : 0.0;
            }
            mbar_wait(&redbar[xs], par);
#pragma unroll
            for (int i = 0; i < NBL; ++i) {
                const int b = lane + 32 * i;
                double sum = 0.0;
                if (b < BT) {
#pragma unroll
                    for (int ww = 0; ww < kV2A; ++ww) sum += red[(xs * kV2A + ww) * BT + b];
                }
                m[i] = sum;
            }
            if (C > 1) {
                if (lane == 0) mbar_arrive_expect_tx(&xbar[xs], C * BT * 8u);
                __syncwarp();
                const uint32_t my_bar = smem_u32(&xbar[xs]);
#pragma unroll
                for (int i = 0; i < NBL; ++i) {
                    const int b = lane + 32 * i;
                    if (b < BT) {
                        const uint32_t my_slot = smem_u32(&xbuf[(xs * C + q) * BT + b]);
                        for (uint32_t d = 0; d < C; ++d) st_async_f64(mapa(my_slot, d), m[i], mapa(my_bar, d));
                    }
                }
                mbar_wait_cluster(&xbar[xs], par);
#pragma unroll
                for (int i = 0; i < NBL; ++i) {
                    const int b = lane + 32 * i;
                    double sum = 0.0;
                    if (b < BT)
                        for (uint32_t d = 0; d < C; ++d) sum += xbuf[(xs * C + d) * BT + b];   // fixed rank order: bit-identical m in every CTA
                    m[i] = sum;
                }
            }
            double r[NBL];
#pragma unroll
            for (int i = 0; i < NBL; ++i) {
                const int b = lane + 32 * i;
                const int64_t bin = tile * BT + b;
                r[i] = 0.0;
                if (b < BT) {
                    if (bin < p.nb) {
                        const double mc = (m[i] < p.eps) ? p.eps : m[i];  // NaN-propagating max
                        r[i] = 1.0 - n_obs[i] / mc;
                    }
                    if (WANT_G) rbuf[xs * BT + b] = r[i];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rbar[xs]);   // B warps: r is published; A warps: red[xs] may be reused
            if (q == 0) {
#pragma unroll
                for (int i = 0; i < NBL; ++i) {
                    const int b = lane + 32 * i;
                    const int64_t bin = tile * BT + b;
                    const bool live = b < BT && bin < p.nb;
                    if (live) {
                        p.composite[bin] = m[i];
                        if (p.residual) p.residual[bin] = r[i];
                    }
                    if (BT >= 32) {
                        if (live) lacc += poisson_term2(m[i], n_obs[i], p.eps);
                    } else {
                        // park this tile's BT pairs in lanes [qcount*BT, (qcount+1)*BT)
                        const int src = lane - qcount * BT;
                        const double vm = __shfl_sync(0xffffffffu, m[i], src & 31);
                        const double vn = __shfl_sync(0xffffffffu, n_obs[i], src & 31);
                        const bool vl = __shfl_sync(0xffffffffu, live ? 1 : 0, src & 31) != 0;
                        if (src >= 0 && src < BT) { qm = vm; qn = vn; qvalid = vl; }
                    }
                }
                if (BT < 32 && ++qcount == QT) {
                    if (qvalid) lacc += poisson_term2(qm, qn, p.eps);
                    qcount = 0;
                    qvalid = false;
                }
            }
        }
        if (q == 0) {
            if (BT < 32 && qcount > 0 && qvalid) lacc += poisson_term2(qm, qn, p.eps);   // the partly filled queue
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) lacc += __shfl_xor_sync(0xffffffffu, lacc, off);
            if (lane == 0) p.lpart[cl] = lacc;
            if (lane == 0 && cl == 0 && p.epoch_ptr) *p.epoch_ptr = *p.epoch_ptr + 1ull;
        }
    }
    // no CTA may exit while a peer can still address its shared memory
    __syncwarp();
    cluster_arrive();
    cluster_wait();
}

}  // namespace sfh
