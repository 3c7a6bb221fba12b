// sfh_fused_pipe.cuh -- K4 with TWO tiles in flight (opt-in: sfh_opts.variant = 3; built in round 1 after the GPU budget was
// spent, so it is NOT the default and has not been measured yet).
//
// sfh_fused.cuh runs pass A -> exchange -> pass B per tile; the cluster exchange costs ~0.8-1.1 us of a ~4.4 us tile during
// which the consumer warps idle (profiles/r1_experiments.md; DESIGN.md section 7 item 1: wide-T F32 stacks sit at ~80 % of
// the HBM roofline because of it).  Here the tile's partial composite is POSTED (st.async to every CTA of the cluster), then
// pass B of the PREVIOUS tile runs from the ring while those partials are in flight, and only then the exchange is waited
// for.  The previous tile's stages therefore stay resident one tile longer: the ring must hold two tiles (ring / G >=
// 2 nst + 1, enforced by choose_config) and the residual buffer is double-buffered by tile parity.  Everything else --
// producers, TMA boxes, fixed-order sums, one store per (cluster, template), bitwise determinism -- is sfh_fused.cuh's;
// the two files are to be merged into one template once this variant has been measured.
#pragma once
#include "sfh_fused.cuh"

namespace sfh {

template <typename S, int BT, int NW>
__global__ void __launch_bounds__((NW + kProducerWarps) * 32, (NW <= 8) ? 2 : 1)
sfh_fg_fused_pipe_kernel(const __grid_constant__ CUtensorMap tmap_full /* box = one whole stage (G chunks) */,
                    const __grid_constant__ CUtensorMap tmap_tail /* box = the tile's last, shorter stage */,
                    const FusedParams p) {
    constexpr bool WANT_G = true, RT = false;
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
    const FusedSmem L = FusedSmem::make(p.ring, BT, (int)C, kt * RPC, NW, G, 1);

    double *red = reinterpret_cast<double *>(smem + L.red_off);    // [NW][BT]
    double *xbuf = reinterpret_cast<double *>(smem + L.xbuf_off);  // [2][C][BT]
    double *rbuf = reinterpret_cast<double *>(smem + L.rbuf_off);  // [2][BT], by tile parity
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
        int ss_prev = 0;          // first ring stage of the previous tile (still resident: its pass B has not run yet)
        bool have_prev = false;
        // pass B of one tile: gradient partials from the bytes still sitting in the ring; releases the stages as it goes
        auto pass_b = [&](int sb, const double *rb) {
            double r[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) r[e] = rb[bl * VEC + e];
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                if (s < nst) {
                    vec16 v[G];
                    const uint32_t sbase = ring_base + (uint32_t)(sb * G) * kChunkBytes + lane_off;
#pragma unroll
                    for (int u = 0; u < G; ++u)
                        if (s * G + u < kt) v[u] = lds128(sbase + (uint32_t)u * kChunkBytes);
#pragma unroll
                    for (int u = 0; u < G; ++u) {
                        if (s * G + u < kt) {
                            double m[VEC];
                            unpack<S>(v[u], m);
                            double g = gacc[s * G + u];
#pragma unroll
                            for (int e = 0; e < VEC; ++e) g = fma(m[e], r[e], g);
                            gacc[s * G + u] = g;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[sb]);
                    if (++sb == NS) sb = 0;
                }
            }
        };
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

            // ---- exchange, part 1: post this CTA's partial to every CTA of the cluster (no wait yet) ----
            const bool xwarp = warp * 32 < BT, active = tid < BT;
            if (active) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < kConsumerWarps; ++w) sum += red[w * BT + tid];
                if (tid == 0) mbar_arrive_expect_tx(&xbar[par], C * BT * 8u);
                const uint32_t my_slot = smem_u32(&xbuf[(par * C + q) * BT + tid]);
                const uint32_t my_bar = smem_u32(&xbar[par]);
                for (uint32_t d = 0; d < C; ++d) st_async_f64(mapa(my_slot, d), sum, mapa(my_bar, d));
            }
            // ---- pass B of the PREVIOUS tile runs while the partials are in flight (the ~1 us exchange bubble) ----
            if (have_prev) pass_b(ss_prev, rbuf + (par ^ 1u) * BT);
            // ---- exchange, part 2: fixed-order sum; residual of THIS tile into rbuf[par] ----
            if (xwarp) {
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
                    rbuf[par * BT + tid] = r;
                }
            }
            // rbuf[par] complete, red[] free for the next tile's pass A
            named_bar_sync<1, kConsumerThreads>();
            ss_prev = ssA;
            have_prev = true;
        }
        if (have_prev) pass_b(ss_prev, rbuf + ((it - 1u) & 1u) * BT);   // the last tile

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
