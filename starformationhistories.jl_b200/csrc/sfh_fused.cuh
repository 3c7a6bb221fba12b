// sfh_fused.cuh -- K4: the fused composite -> residual -> transposed-gradient kernel (sm_100a).
//
// Replaces, in ONE pass over the template stack, the reference's
//     composite!            src/fitting/fitting_base.jl:55-65   (gemv 'N', reads M)
//     grad-loglikelihood!   src/fitting/fitting_base.jl:265-285 (residual + gemv 'T', reads M again)
// as sequenced by fg! (src/fitting/solvers.jl:20-38).  The O(Nb) Poisson term of
// loglikelihood (fitting_base.jl:84-96) is evaluated by the finalize kernel from the composite
// vector this kernel writes (sfh_small.cuh).
//
// Decomposition
//   stack M: column-major Nb x T, leading dimension ld (stack_models layout, padded).
//   tile      = BT consecutive bins x ALL T templates; owned by one thread-block CLUSTER.
//   CTA q of the cluster holds templates [q*KT*RPC, (q+1)*KT*RPC) of the tile in shared memory,
//   brought in by TMA as KT chunks (boxes of BT bins x RPC templates = 8 KB each) through an
//   mbarrier ring of R slots that is deeper than one tile, so the next tile streams in while this
//   one is being used.
//   pass A   every consumer lane owns 16 B of each chunk (VEC bins of one template) and FMAs it
//            into its composite partials;  warp shuffles + one smem stage give the CTA partial;
//            the C CTA partials are exchanged through DSMEM with st.async (data + mbarrier
//            complete_tx in one op) and summed in fixed rank order -> every CTA holds bit-identical m.
//   residual r = 1 - n/max(m,eps)  (fitting_base.jl:277-279)
//   pass B   the same lanes re-read the same 16 B from SHARED memory (not HBM, not L2) and
//            accumulate M*r into per-lane gradient partials that live in registers across ALL tiles
//            of the CTA; slots are released to the TMA producer as pass B walks them.
//   end      lanes sharing a template combine by shuffle; one plain store per (cluster, template)
//            into gpart[n_clusters][T].  No atomics anywhere => bitwise run-to-run determinism.
#pragma once
#include "sfh_ptx.cuh"

namespace sfh {

// NW consumer warps (+1 TMA producer warp) per CTA.  NW = 16: one CTA per SM; NW = 8: two CTAs per SM, so one
// CTA's exchange/residual latency is hidden behind the other CTA's streaming passes.
constexpr int kKMax = 20;       // max chunks per CTA per tile (per-lane register array gacc[])
// per-variant limit: the 12-warp register-tile variant has 128 registers/thread => 12 chunks (4+2 regs each)
__host__ __device__ constexpr int kmax_for(int nw, bool rt) { return (rt && nw == 12) ? 12 : kKMax; }
// chunks per pipeline STAGE: one mbarrier wait / release per stage instead of per chunk (an already-complete
// mbarrier.try_wait still costs ~90 cycles; with few warps per SM that latency, paid per chunk, made the
// consumers -- not HBM -- the bottleneck).  kKMax and every kmax_for() are multiples of it.
#ifndef SFH_PRODUCERS
#define SFH_PRODUCERS 2
#endif
#ifndef SFH_STAGE_SMEM
#define SFH_STAGE_SMEM 2
#endif
__host__ __device__ constexpr int stage_chunks_for(bool rt) { return rt ? 4 : SFH_STAGE_SMEM; }
// (one op per 2-chunk stage + 2 producers: 194 us; 1-chunk stages + 1 producer: 206 us; see r1_experiments.md)
constexpr int kMaxCluster = 16;
__host__ __device__ constexpr int chunk_bytes(int nw) { return nw * 32 * 16; }  // one 16-byte vector per consumer lane
// TMA producer warps per CTA.  A single elected thread needs ~250 cycles per chunk (mbarrier try_wait on the empty
// slot ~100, expect_tx, bulk-tensor issue): at 4 KB chunks that caps a CTA at ~31 GB/s (profiles/probes/bw_probe.cu
// reproduces it: 4 KB x 26-slot ring, 1 producer thread, 1 CTA/SM -> 4.5 TB/s; 16 KB chunks -> 7.0 TB/s).
// Two producers interleave the chunk sequence; 10 (or 18) warps still fit the 96-register budget of 5 warps/scheduler.
constexpr int kProducerWarps = SFH_PRODUCERS;  // see profiles/r1_experiments.md for the producers x stage-size matrix

// a pipeline stage (G chunks = G*RPC templates of the tile) is ONE bulk-tensor op
constexpr int kMaxStage = 4;

struct FusedParams {
    int64_t nb;          // bins in this shard
    int64_t nt;          // templates
    int32_t kt;          // chunks per CTA per tile (<= kKMax)
    int32_t ring;        // ring chunk-slots: a multiple of the stage size
    int32_t n_tiles;     // ceil(nb / BT)
    int32_t evict_first; // use an L2 evict_first policy on the stack loads
    int32_t l2_prefetch; // tiles of look-ahead for cp.async.bulk.prefetch.tensor (0 = off)
    int32_t panel;       // 1: the stack is stored as bin-major panels [tile][T][BT] and the tensor map is 3-D
    double eps;          // clamp (fitting_base.jl:90,277)
    const double *coeffs;   // [nt]
    const double *data;     // [nb] (converted to double at upload)
    double *composite;      // [nb] out: M*coeffs (unclamped)
    double *residual;       // [nb] out (nullable): 1 - n/max(m,eps)
    double *gpart;          // [n_clusters][gstride] out
    int64_t gstride;
};

template <typename S, int BT, int NW>
struct FusedCfg {
    static constexpr int VEC = 16 / sizeof(S);  // elements per 16-byte lane vector
    static constexpr int LPR = BT / VEC;        // lanes per template row
    static constexpr int RPW = 32 / LPR;        // template rows per warp per chunk
    static constexpr int RPC = RPW * NW;        // template rows per chunk
    static_assert(BT % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "bad tile");
    static_assert(RPC * BT * sizeof(S) == chunk_bytes(NW), "chunk = one 16-byte vector per consumer lane");
    static_assert(RPC <= 256, "TMA box dimension limit");
};

// dynamic shared memory carve-up (bytes), shared by host and device
struct FusedSmem {
    uint32_t ring_off, red_off, xbuf_off, rbuf_off, cs_off, bar_off, total;
    // cs_elems = kt * RPC: this CTA's slice of the coefficient vector (kept in smem, not registers:
    // 17 warps put 5 warps on one SM sub-partition => 96 registers/thread, too few for c[] + gacc[])
    __host__ __device__ static FusedSmem make(int ring, int bt, int cluster, int cs_elems, int nw, int g) {
        FusedSmem s;
        s.ring_off = 0;
        s.red_off = ring * chunk_bytes(nw);
        s.xbuf_off = s.red_off + nw * bt * 8;
        s.rbuf_off = s.xbuf_off + 2 * cluster * bt * 8;
        s.cs_off = s.rbuf_off + bt * 8;
        s.bar_off = s.cs_off + cs_elems * 8;
        s.total = s.bar_off + (2 * (ring / g) + 2) * 8;
        return s;
    }
};

template <typename S>
__device__ __forceinline__ void unpack(const vec16 &v, double (&out)[16 / sizeof(S)]);
template <>
__device__ __forceinline__ void unpack<double>(const vec16 &v, double (&out)[2]) {
    out[0] = __hiloint2double(v.y, v.x);
    out[1] = __hiloint2double(v.w, v.z);
}
template <>
__device__ __forceinline__ void unpack<float>(const vec16 &v, double (&out)[4]) {
    out[0] = (double)__uint_as_float(v.x);
    out[1] = (double)__uint_as_float(v.y);
    out[2] = (double)__uint_as_float(v.z);
    out[3] = (double)__uint_as_float(v.w);
}

// RT ("register tile"): pass A keeps its 16-byte vectors in registers for pass B and releases the smem stage at
// once, so shared memory is a pure streaming ring (every slot in flight) and the HBM stream is decoupled from
// the A -> exchange -> B dependency.  Needs 4*KT more registers per lane => NW <= 12, one CTA per SM.
template <typename S, int BT, int NW, bool WANT_G, bool RT>
__global__ void __launch_bounds__((NW + kProducerWarps) * 32, (NW <= 8 && !RT) ? 2 : 1)
sfh_fg_fused_kernel(const __grid_constant__ CUtensorMap tmap_full /* box = one whole stage (G chunks) */,
                    const __grid_constant__ CUtensorMap tmap_tail /* box = the tile's last, shorter stage */,
                    const FusedParams p) {
    using Cfg = FusedCfg<S, BT, NW>;
    constexpr int VEC = Cfg::VEC, LPR = Cfg::LPR, RPW = Cfg::RPW, RPC = Cfg::RPC;
    constexpr int kConsumerWarps = NW, kConsumerThreads = NW * 32, kFusedThreads = (NW + kProducerWarps) * 32;
    constexpr uint32_t kChunkBytes = chunk_bytes(NW);
    constexpr int KMAX = kmax_for(NW, RT);
    constexpr int G = stage_chunks_for(RT);
    constexpr int SMAX = KMAX / G;  // stages per tile at most
    static_assert(KMAX % G == 0 && G <= kMaxStage, "KMAX must be a multiple of the stage size");
    constexpr bool kOneOp = (G * RPC <= 256);  // a whole stage fits one TMA box (box dimensions are limited to 256)

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t q = cluster_ctarank();
    const uint32_t C = cluster_nctarank();
    const uint32_t cl = cluster_id_x();
    const uint32_t ncl = cluster_nid_x();
    const int kt = p.kt;
    const int NS = p.ring / G;          // stage slots in the ring
    const int nst = (kt + G - 1) / G;   // stages per tile
    const FusedSmem L = FusedSmem::make(p.ring, BT, (int)C, kt * RPC, NW, G);

    double *red = reinterpret_cast<double *>(smem + L.red_off);    // [NW][BT]
    double *xbuf = reinterpret_cast<double *>(smem + L.xbuf_off);  // [2][C][BT]
    double *rbuf = reinterpret_cast<double *>(smem + L.rbuf_off);  // [BT]
    double *cs = reinterpret_cast<double *>(smem + L.cs_off);      // [kt*RPC]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint64_t *empty = full + NS;
    uint64_t *xbar = empty + NS;  // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kConsumerWarps);
        }
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        fence_mbar_init();
    }
    if (warp == kConsumerWarps && lane == 0) {
        prefetch_tensormap(&tmap_full);
        prefetch_tensormap(&tmap_tail);
    }
    // PDL: barrier set-up above overlaps the previous kernel; its outputs (coeffs) are read only below
    griddep_wait();
    for (int i = tid; i < kt * RPC; i += kFusedThreads) {
        const int64_t j = (int64_t)q * kt * RPC + i;
        cs[i] = (j < p.nt) ? __ldg(p.coeffs + j) : 0.0;
    }
    __syncthreads();
    // every CTA's barriers must be initialised before any peer signals them
    cluster_arrive();
    cluster_wait();

    if (warp >= kConsumerWarps) {
        // ================= TMA producers: warp pw issues the stages gs = pw, pw + NP, pw + 2 NP, ... =================
        if (lane == 0) {
            const int pw = warp - kConsumerWarps;
            const uint64_t pol = l2_policy_evict_first();
            const int32_t t0 = (int32_t)(q * (uint32_t)(kt * RPC));
            const int pf = p.l2_prefetch * (int)ncl;  // look-ahead in tiles of THIS cluster's sequence
            const int my_tiles = (p.n_tiles > (int)cl) ? (p.n_tiles - (int)cl + (int)ncl - 1) / (int)ncl : 0;
            // stages this CTA consumes, in order: (it, s) advanced incrementally -- a 64-bit div/mod per stage here
            // costs the single issuing thread ~250 cycles and showed up as a 1.8x slowdown of the whole kernel
            int ss = pw % NS;
            uint32_t round = (uint32_t)(pw / NS);
            int it = pw / nst, s = pw % nst;
            for (; it < my_tiles;) {
                const int tile = (int)cl + it * (int)ncl;
                const int cnt = (kt - s * G < G) ? (kt - s * G) : G;
                // (experiment knob) pull the same chunks of a LATER tile into L2: not gated by a free smem slot
                // box of cnt*RPC templates: the whole stage in ONE op when it fits a TMA box, else one op per chunk
                const int nops = kOneOp ? 1 : cnt;
                // (two __grid_constant__ maps selected by a ternary: indexing an array of maps dynamically would make
                //  nvcc copy it to local memory, and maps fetched from global memory halve the issue rate)
                const CUtensorMap *tmap = (kOneOp && cnt != G) ? &tmap_tail : &tmap_full;
                if (round > 0) mbar_wait(&empty[ss], (round - 1) & 1u);
                mbar_arrive_expect_tx(&full[ss], (uint32_t)cnt * kChunkBytes);
                for (int u = 0; u < nops; ++u) {
                    const int32_t tc = t0 + (s * G + u) * RPC;
                    void *dst = smem + L.ring_off + (uint32_t)(ss * G + u) * kChunkBytes;
                    if (pf > 0 && tile + pf < p.n_tiles && tc < (int32_t)p.nt) tma_prefetch_l2_2d(tmap, (tile + pf) * BT, tc);
                    if (p.panel) {
                        if (p.evict_first)
                            tma_load_3d_hint(dst, tmap, 0, tc, tile, &full[ss], pol);
                        else
                            tma_load_3d(dst, tmap, 0, tc, tile, &full[ss]);
                    } else if (p.evict_first)
                        tma_load_2d_hint(dst, tmap, tile * BT, tc, &full[ss], pol);
                    else
                        tma_load_2d(dst, tmap, tile * BT, tc, &full[ss]);
                }
                ss += kProducerWarps;
                if (ss >= NS) { ss -= NS; ++round; }
                s += kProducerWarps;
                while (s >= nst) { s -= nst; ++it; }
            }
        }
    } else {
        // ================= consumers: NW warps, one 16-byte vector per lane per chunk =========
        const int bl = lane % LPR;  // which VEC-bin group of the tile this lane owns
        const int rw = lane / LPR;  // which template row of the warp's RPW rows
        const uint32_t lane_off = (uint32_t)tid * 16u;
        const uint32_t ring_base = smem_u32(smem + L.ring_off);

        // the lane's templates: j(k) = q*kt*RPC + k*RPC + warp*RPW + rw  (fixed for the whole kernel)
        const int64_t j0 = (int64_t)q * kt * RPC + warp * RPW + rw;
        const double *cs_lane = cs + warp * RPW + rw;  // + k*RPC
        double gacc[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) gacc[k] = 0.0;

        vec16 tile_regs[RT ? KMAX : 1];  // RT: this lane's 16 bytes of every chunk of the current tile
        int ss = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (int tile = (int)cl; tile < p.n_tiles; tile += (int)ncl, ++it) {
            const uint32_t par = it & 1u;
            const int ssA = ss;

            // the tile's observed counts: issued now, consumed after the exchange (off the critical path)
            double n_obs = 0.0;
            if (tid < BT && (int64_t)tile * BT + tid < p.nb) n_obs = __ldg(p.data + (int64_t)tile * BT + tid);

            // ---- pass A: composite partials for this lane's VEC bins over its templates ----
            double acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                if (s < nst) {
                    mbar_wait(&full[ss], phase);
                    vec16 v[G];
                    double ck[G];
                    const uint32_t sbase = ring_base + (uint32_t)(ss * G) * kChunkBytes + lane_off;
#pragma unroll
                    for (int u = 0; u < G; ++u) {  // G independent shared-memory loads in flight
                        if (s * G + u < kt) {
                            v[u] = lds128(sbase + (uint32_t)u * kChunkBytes);
                            ck[u] = cs_lane[(s * G + u) * RPC];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < G; ++u) {
                        if (s * G + u < kt) {
                            double m[VEC];
                            unpack<S>(v[u], m);
#pragma unroll
                            for (int e = 0; e < VEC; ++e) acc[e] = fma(m[e], ck[u], acc[e]);
                            if (RT) tile_regs[RT ? s * G + u : 0] = v[u];
                        }
                    }
                    if (!WANT_G || RT) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[ss]);
                    }
                    if (++ss == NS) { ss = 0; phase ^= 1u; }
                }
            }
            // lanes that own the same bins (different rows) combine
#pragma unroll
            for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
            }
            if (rw == 0) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) red[warp * BT + bl * VEC + e] = acc[e];
            }
            named_bar_sync<1, kConsumerThreads>();

            // ---- exchange: CTA partial -> all CTAs of the cluster; fixed-order sum; residual ----
            if (warp * 32 < BT) {
                const bool active = tid < BT;
                if (active) {
                    double sum = 0.0;
#pragma unroll
                    for (int w = 0; w < kConsumerWarps; ++w) sum += red[w * BT + tid];
                    if (tid == 0) mbar_arrive_expect_tx(&xbar[par], C * BT * 8u);
                    const uint32_t my_slot = smem_u32(&xbuf[(par * C + q) * BT + tid]);
                    const uint32_t my_bar = smem_u32(&xbar[par]);
                    for (uint32_t d = 0; d < C; ++d) st_async_f64(mapa(my_slot, d), sum, mapa(my_bar, d));
                }
                mbar_wait_cluster(&xbar[par], (it >> 1) & 1u);
                if (active) {
                    double m = 0.0;
                    for (uint32_t d = 0; d < C; ++d) m += xbuf[(par * C + d) * BT + tid];
                    const int64_t bin = (int64_t)tile * BT + tid;
                    double r = 0.0;
                    if (bin < p.nb) {
                        const double n = n_obs;
                        const double mc = (m < p.eps) ? p.eps : m;  // NaN-propagating max
                        r = 1.0 - n / mc;
                        if (q == 0) {
                            p.composite[bin] = m;
                            if (p.residual) p.residual[bin] = r;
                        }
                    }
                    rbuf[tid] = r;
                }
            }
            named_bar_sync<1, kConsumerThreads>();

            // ---- pass B: gradient partials from the SAME bytes (shared memory, or registers if RT) ----
            if (WANT_G) {
                double r[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) r[e] = rbuf[bl * VEC + e];
                int sb = ssA;
#pragma unroll
                for (int s = 0; s < SMAX; ++s) {
                    if (s < nst) {
                        vec16 v[G];
                        if (!RT) {
                            const uint32_t sbase = ring_base + (uint32_t)(sb * G) * kChunkBytes + lane_off;
#pragma unroll
                            for (int u = 0; u < G; ++u)
                                if (s * G + u < kt) v[u] = lds128(sbase + (uint32_t)u * kChunkBytes);
                        }
#pragma unroll
                        for (int u = 0; u < G; ++u) {
                            if (s * G + u < kt) {
                                const vec16 vv = RT ? tile_regs[RT ? s * G + u : 0] : v[u];
                                double m[VEC];
                                unpack<S>(vv, m);
                                double g = gacc[s * G + u];
#pragma unroll
                                for (int e = 0; e < VEC; ++e) g = fma(m[e], r[e], g);
                                gacc[s * G + u] = g;
                            }
                        }
                        if (!RT) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&empty[sb]);
                            if (++sb == NS) sb = 0;
                        }
                    }
                }
            }
        }

        // ---- end of kernel: one store per (cluster, template) ----
        if (WANT_G) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < kt) {
                    double g = gacc[k];
#pragma unroll
                    for (int off = 1; off < LPR; off <<= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
                    const int64_t j = j0 + (int64_t)k * RPC;
                    if (bl == 0 && j < p.nt) p.gpart[(int64_t)cl * p.gstride + j] = g;
                }
            }
        }
    }
    // no CTA may exit while a peer can still address its shared memory
    __syncwarp();
    cluster_arrive();
    cluster_wait();
}

}  // namespace sfh
