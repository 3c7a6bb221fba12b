// sfh_small.cuh -- the O(Nb) / O(T) kernels around K4, the two-pass (unfused) kernels that give
// signature-level parity for direct calls of composite! / loglikelihood / grad-loglikelihood!,
// and the on-device synthetic generators used by the large benchmark configurations.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sfh_ptx.cuh"

namespace sfh {

// ------------------------------------------------------------------------------------------
// Device layout of the template stack.  The HOST-visible layout is always stack_models' column-major Nb x T
// (src/fitting/utilities.jl:12-13); on the device the fused path stores it as BIN-MAJOR PANELS of `bt` bins:
//     element (i, j)  ->  ((i / bt) * nt + j) * bt + (i % bt)
// so that everything one cluster needs for a bin tile is ONE contiguous nt*bt block (sequential DRAM pages, one
// TLB entry per tile instead of one per template: 2.2x on the 40 GB stack, +7 % at config 3 -- SURVEY.md section 7
// option (c)).  panel == 0 keeps plain column-major with leading dimension ld (two-pass / non-sm_100 path).
// ------------------------------------------------------------------------------------------
struct StackLayout {
    int64_t ld, nt, rows;
    int32_t bt_shift, panel;
    __host__ __device__ __forceinline__ int64_t off(int64_t i, int64_t j) const {
        return panel ? ((((i >> bt_shift) * nt + j) << bt_shift) + (i & ((1 << bt_shift) - 1))) : (i + j * ld);
    }
    __host__ __device__ __forceinline__ int64_t alloc_elems() const {
        const int64_t bt = (int64_t)1 << bt_shift;
        return panel ? ((rows + bt - 1) >> bt_shift) * nt * bt : ld * nt;
    }
};

// ------------------------------------------------------------------------------------------
// deterministic block reduction (fixed shuffle tree + fixed smem order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sh /*[NT/32]*/) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = (lane < NT / 32) ? sh[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;  // valid in warp 0
}

// Poisson log-likelihood-ratio term, fitting_base.jl:90-92.  `ifelse` semantics: select.
__device__ __forceinline__ double poisson_term(double m, double n, double eps) {
    const double mc = (m < eps) ? eps : m;  // NaN propagates like Julia's scalar max
    return (n > 0.0) ? (n - mc - n * log(n / mc)) : -mc;
}

// ------------------------------------------------------------------------------------------
// Chain rule of the hierarchical fg! (mzr.jl:124-210 / amr.jl:118-169) inside the finalize step: everything in it that does not
// depend on the gradient (the per-template factors W1..W4) is prepared by sfh_hier_prologue2_kernel while the variables are
// being turned into coefficients; sfh_finalize_hier_kernel (one block per age group) reduces the gradient of the group's
// templates, multiplies by the factors and leaves four sums per age; its last block finishes with 60-element sums and scans.
// Round 1 ran a fourth, single-block epilogue launch for this (22 us).
// ------------------------------------------------------------------------------------------
constexpr int kHierTailAges = 256;   // ages the folded path stages in shared memory (more: the separate epilogue kernel)
struct HierTail {
    int32_t on;          // 0 = plain fg!
    int32_t kind, nj, want_G;
    uint8_t free_mask[4];
    const double *W;     // [4][nt], AGE-GROUP order (W[f][g], g as in gmem): the factors of  sum_k fullG_jk * (...)  for
                         // R_j (cross-age), R_j (same age), mu_j, sigma
    double *sums;        // [4][nj] scratch: the four per-age sums, one block each
    const double *gA, *gB;   // [nj] d mu_j / d alpha, d mu_j / d beta
    const int32_t *gptr, *gmem, *sidx;
    int64_t nt;
    double *out;         // device [1 + nj + 3]: -logL (guarded), G
    double *out_host;    // nullable mapped pinned copy
    long long *dbg;      // nullable: see FinalizeParams::dbg
};

// ------------------------------------------------------------------------------------------
// finalize: logL = sum_i term(m_i, n_i)   (loglikelihood, fitting_base.jl:84-96, raw sum: the
// `== 0 -> -Inf` guard of :95 is applied by the caller AFTER any cross-GPU all-reduce)
// and G_j = sum over clusters of gpart[cl][j].  out = [logL, G_0..G_{T-1}].
// ------------------------------------------------------------------------------------------
struct FinalizeParams {
    int64_t nb, nt, gstride;
    int32_t n_clusters, want_G;
    int32_t nblk_logl;  // blocks that own bins: depends on nb only, so logL is bitwise identical with and without G
    double eps;
    const double *composite, *data, *gpart;
    double *out;       // device [1 + nt]
    double *out_host;  // nullable: mapped pinned host copy of the same (saves the D2H memcpy of the sync API)
    double *lpart;     // [gridDim.x] per-block logL partials
    const double *lpart_in;  // nullable: the fused kernel (v2) already summed the Poisson terms per cluster -> [n_lpart_in]
    int32_t n_lpart_in;
    unsigned int *ticket;
    // K7 v2: one-shot all-reduce over NVLink peer memory, fused into this kernel's tail (nullptr = off).
    // peers[r] = rank r's inbox: [2 parities][nranks][vlen] 16-byte self-validating packets (st_packet).
    double *const *peers;
    int32_t nranks, rank;
    int64_t vlen;
    int32_t epoch_from_fused;       // 1: the stream kernel of this evaluation has already bumped *epoch_ptr (Fused2Params::epoch_ptr):
                                    // the epoch is read as it is, and the flat kernel needs neither ticket nor last block
    unsigned long long *epoch_ptr;  // device: evaluations exchanged so far on this context (read and bumped by the last
                                    // block only, so a captured CUDA graph of the evaluation can be replayed)
    HierTail hier;                  // hierarchical chain rule on the (all-reduced) gradient, by the last block
    // Completion without a stream synchronisation: when pkt_host is set, every result is ALSO stored into the caller's mapped
    // pinned buffer as a self-validating packet (st_packet) carrying the epoch the host wrote next to its inputs; the host polls
    // the packets (sfh_api.cu: wait_packets) instead of waiting for the stream to drain.  pkt_epoch = device copy of that epoch
    // (written by the upload kernel / the hierarchical prologue of the same evaluation).
    void *pkt_host;
    const unsigned long long *pkt_epoch;
    int32_t pdl_early;              // release the dependent launch at kernel start
    long long *dbg;                 // nullable (SFH_DEBUG_FINALIZE=1): clock64() of the last block's thread 0 at its milestones
};

// One value of the exchange = one 16-byte packet {lo32, epoch32, hi32, epoch32}: each 8-byte half carries its own flag, so the
// packet validates itself however the store is split on the way (8-byte aligned stores are single transactions) -- no fence and
// no separate flag store between data and signal, which is what made the first two versions cost 17-18 us per step at 2 GPUs
// against NCCL's 12 (profiles/r2_experiments.md).
__device__ __forceinline__ void st_packet(void *dst, double v, uint32_t ep) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((uint32_t)b), "r"(ep), "r"((uint32_t)(b >> 32)), "r"(ep)
                 : "memory");
}
__device__ __forceinline__ double ld_packet_wait(const void *src, uint32_t ep) {
    uint32_t lo, f0, hi, f1;
    do {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src) : "memory");
    } while (f0 != ep || f1 != ep);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
constexpr int kFinalizeThreads = 256;
constexpr int kFinE = 4;      // entries a warp of the sharded finalize kernel has in flight (pushed before any is polled)
constexpr int kPktG0 = 8;     // host packets of the flat call: [0] = logL, [kPktG0 + j] = G_j -- a block's 8 packets start on a 128-byte boundary
constexpr int kMaxExchangeRanks = 32;   // ranks of the one-shot exchange (sfh_comm_p2p_init / sfh_group_create refuse more)

enum { MH_POWERLAW_MZR = 0, MH_LINEAR_AMR = 1, MH_LOG_AMR = 2 };

// Everything here is latency, not bandwidth (~1 MB of partials): the shape is chosen so that no thread ever waits on more than
// ~2 dependent L2 round trips.  Measured alternatives (ncu launch lists under profiles/): 296 blocks with a per-thread serial loop
// over the cluster partials 13 us; one 8-CTA cluster with a DSMEM reduction 27 us (too few threads).
//
// Multi-GPU (p.peers): the exchange happens PER GRADIENT ENTRY in the warp that produced it -- the warp stores its entry as a
// self-validating packet into slot [parity][rank] of every rank's inbox, polls the nranks packets of the same entry in its own
// inbox and sums them in rank order (identical order on every rank => bit-identical results everywhere).  2400 warps exchange
// in parallel: one NVLink round trip for the whole vector, no fence, no flag, no extra kernel, no block that serialises the sum
// (three earlier versions did, at 16-18 us per step on 2 GPUs; profiles/r2_experiments.md section 3).  logL travels the same way
// from block 0.  Block b handles the same entries on every rank and pushes before it polls, so the ranks' grids cannot wait on
// each other in a cycle even if a grid were larger than what is co-resident.  The two parities of the inbox suffice: a rank can
// be at most one evaluation ahead of a peer, because finishing evaluation k+1 needs the peer's packets of k+1, which the peer
// sends only after it has finished reading those of k (stream order).

// raw logL of this shard -> all-reduced over the ranks (same packet protocol, slot 0 of the vector).  With the stream kernel the
// shard's logL exists when the finalize kernel STARTS (per-cluster Poisson partials), so block 0 pushes it at once and collects
// the sum after its own gradient entries (the hierarchical kernel: in its last block): the logL round trip overlaps the gradient's.
__device__ __forceinline__ void push_logl(const FinalizeParams &p, double all, int lane, int64_t par, uint32_t ep32) {
    if (lane < p.nranks) st_packet(reinterpret_cast<uint4 *>(p.peers[lane]) + (par * p.nranks + p.rank) * p.vlen, all, ep32);
}
__device__ __forceinline__ double poll_logl(const FinalizeParams &p, int lane, int64_t par, uint32_t ep32) {
    const uint4 *inbox = reinterpret_cast<const uint4 *>(p.peers[p.rank]) + par * p.nranks * p.vlen;
    const double v = (lane < p.nranks) ? ld_packet_wait(inbox + (int64_t)lane * p.vlen, ep32) : 0.0;
    double t = 0.0;
#pragma unroll 1
    for (int r = 0; r < p.nranks; ++r) t += __shfl_sync(0xffffffffu, v, r);
    return t;
}
// block 0, at kernel start: the fused kernel's per-cluster Poisson partials -> this shard's logL (fixed order); returns it in warp 0
__device__ __forceinline__ double early_logl(const FinalizeParams &p, double *sh) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < p.n_lpart_in; b += kFinalizeThreads) acc += __ldcg(p.lpart_in + b);
    return block_sum<kFinalizeThreads>(acc, sh);
}

__global__ void __launch_bounds__(kFinalizeThreads) sfh_finalize_kernel(const FinalizeParams p) {
    __shared__ double sh[kFinalizeThreads / 32];
    __shared__ double shg[2][kFinE][kFinalizeThreads / 32];   // a pass's results, gathered for the host packets
    __shared__ bool last;
    if (p.pdl_early) griddep_launch_dependents();   // the next evaluation's first kernel may become resident (it waits in turn)
    griddep_wait();  // PDL: launched while the fused kernel drains
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t hep = p.pkt_host ? (uint32_t)__ldcg(p.pkt_epoch) : 0u;   // the host's epoch of this evaluation
    uint4 *const pkt = reinterpret_cast<uint4 *>(p.pkt_host);
    // the epoch of THIS evaluation: bumped by the stream kernel before this grid started, or the stored one + 1 (then bumped by
    // the last block only, after every block has read it)
    const unsigned long long epoch = p.peers ? __ldcg(p.epoch_ptr) + (p.epoch_from_fused ? 0ull : 1ull) : 0ull;
    const uint32_t ep32 = (uint32_t)epoch;
    const int64_t par = (int64_t)(epoch & 1ull);
    if (p.lpart_in) {
        // stream kernel: the shard's logL is a sum of per-cluster partials that already exist -- block 0 finishes it right away
        if (blockIdx.x == 0) {
            const double all = early_logl(p, sh);
            if (warp == 0) {
                if (p.peers) push_logl(p, all, lane, par, ep32);
                else if (lane == 0) { p.out[0] = all; if (p.out_host) p.out_host[0] = all; if (pkt) st_packet(pkt, all, hep); }
            }
        }
    } else {
        // logL from the composite: fixed contiguous slice of bins per block, fixed trees => deterministic
        const int64_t per = (p.nb + p.nblk_logl - 1) / p.nblk_logl;
        const int64_t b0 = (int64_t)blockIdx.x * per;
        const int64_t b1 = ((int)blockIdx.x >= p.nblk_logl) ? b0 : ((b0 + per < p.nb) ? b0 + per : p.nb);
        double acc = 0.0;
        for (int64_t i = b0 + threadIdx.x; i < b1; i += kFinalizeThreads)
            acc += poisson_term(__ldcg(p.composite + i), p.data[i], p.eps);
        const double tot = block_sum<kFinalizeThreads>(acc, sh);
        if (threadIdx.x == 0) p.lpart[blockIdx.x] = tot;
    }

    // G_j = sum over clusters: one WARP per template, lanes take clusters lane, lane+32, ... (independent loads).  A block owns 8
    // CONSECUTIVE templates per pass; their results leave for the host together: 8 packets = 128 aligned bytes in ONE store
    // instruction of warp 0, i.e. whole 64-byte lines.  (2400 separate 16-byte stores, four to a line and arriving at different
    // times while the host polls that very line, cost +8 us per call on one of the boxes measured and -4 on another:
    // profiles/r2_experiments.md section 6a.)
    if (p.want_G) {
        constexpr int NW = kFinalizeThreads / 32;
        const int64_t nwarps = (int64_t)gridDim.x * NW;
        const int64_t jb0 = (int64_t)blockIdx.x * NW;
        if (!p.peers) {
            int buf = 0;
            for (int64_t jb = jb0; jb < p.nt; jb += nwarps, buf ^= 1) {
                const int64_t j = jb + warp;
                if (j < p.nt) {
                    double s = 0.0;
                    for (int cl = lane; cl < p.n_clusters; cl += 32) s += __ldcg(p.gpart + (int64_t)cl * p.gstride + j);
                    s = warp_sum(s);   // xor tree: every lane holds the total
                    if (lane == 0) {
                        p.out[1 + j] = s;
                        if (p.out_host) p.out_host[1 + j] = s;
                        shg[buf][0][warp] = s;
                    }
                }
                if (pkt) {
                    __syncthreads();   // (one barrier per pass: the two buffers alternate)
                    if (warp == 0 && lane < NW && jb + lane < p.nt) st_packet(pkt + kPktG0 + jb + lane, shg[buf][0][lane], hep);
                }
            }
        } else {
            // sharded: a warp that owns several entries (T > warps of the grid, e.g. 10^4 templates) pushes ALL of them before it
            // polls any, so their NVLink round trips overlap instead of following one another (config 5 at 8 GPUs: 3 per warp)
            const uint4 *inbox = reinterpret_cast<const uint4 *>(p.peers[p.rank]) + par * p.nranks * p.vlen;
            int buf = 0;
            for (int64_t jb = jb0; jb < p.nt; jb += kFinE * nwarps, buf ^= 1) {
#pragma unroll
                for (int e = 0; e < kFinE; ++e) {
                    const int64_t j = jb + e * nwarps + warp;
                    if (j < p.nt) {
                        double s = 0.0;
                        for (int cl = lane; cl < p.n_clusters; cl += 32) s += __ldcg(p.gpart + (int64_t)cl * p.gstride + j);
                        s = warp_sum(s);
                        if (lane < p.nranks) st_packet(reinterpret_cast<uint4 *>(p.peers[lane]) + (par * p.nranks + p.rank) * p.vlen + 1 + j, s, ep32);
                    }
                }
#pragma unroll
                for (int e = 0; e < kFinE; ++e) {
                    const int64_t j = jb + e * nwarps + warp;
                    if (j < p.nt) {
                        const double v = (lane < p.nranks) ? ld_packet_wait(inbox + (int64_t)lane * p.vlen + 1 + j, ep32) : 0.0;
                        double t = 0.0;
#pragma unroll 1
                        for (int r = 0; r < p.nranks; ++r) t += __shfl_sync(0xffffffffu, v, r);   // rank order: bit-identical everywhere
                        if (lane == 0) {
                            p.out[1 + j] = t;
                            if (p.out_host) p.out_host[1 + j] = t;
                            shg[buf][e][warp] = t;
                        }
                    }
                }
                if (pkt) {
                    __syncthreads();
                    if (warp < kFinE && lane < NW) {   // warp e sends pass e's 8 packets
                        const int64_t j = jb + (int64_t)warp * nwarps + lane;
                        if (j < p.nt) st_packet(pkt + kPktG0 + j, shg[buf][warp][lane], hep);
                    }
                }
            }
        }
    }
    if (p.lpart_in && !p.peers) return;   // single GPU, stream kernel: nothing is left for a last block -- no fence, no ticket
    if (p.lpart_in && p.epoch_from_fused) {
        // sharded, stream kernel: block 0 pushed the shard's logL at kernel start and collects the sum here, after its own gradient
        // entries; the epoch was bumped by the fused kernel -- no fence, no ticket, no last block either
        if (blockIdx.x == 0 && warp == 0) {
            const double all = poll_logl(p, lane, par, ep32);
            if (lane == 0) {
                p.out[0] = all;
                if (p.out_host) p.out_host[0] = all;
                if (pkt) st_packet(pkt, all, hep);
            }
        }
        return;
    }
    // the last block: folds the per-block logL partials (parallel, fixed order) / finishes the logL exchange, bumps the epoch
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    double all = 0.0;
    if (!p.lpart_in) {
        double s = 0.0;
        for (int b = threadIdx.x; b < p.nblk_logl; b += kFinalizeThreads) s += __ldcg(p.lpart + b);
        all = block_sum<kFinalizeThreads>(s, sh);   // valid in warp 0
        if (p.peers && warp == 0) push_logl(p, all, lane, par, ep32);
    }
    if (p.peers && warp == 0) all = poll_logl(p, lane, par, ep32);
    if (threadIdx.x == 0) {
        p.out[0] = all;
        if (p.out_host) p.out_host[0] = all;
        if (pkt) st_packet(pkt, all, hep);
        *p.ticket = 0u;  // re-arm for the next evaluation on this context
        if (p.peers) *p.epoch_ptr = epoch;
    }
}

// The finalize step of the HIERARCHICAL evaluation: block j = age group j.  Its 8 warps each add a fixed eighth of the cluster
// partials for the group's templates (lane = member, so the loads of a warp are one contiguous row of gpart per cluster), warp 0
// combines them in fixed order, exchanges the entries when sharded (lane = member: pushes to and polls from every rank),
// stores G_t, multiplies fullG_t = -G_t by the four chain-rule factors (W in age-group order: coalesced) and reduces over the
// members: four numbers per age.  The last block adds logL, the three parameter sums (mzr.jl:196-208) and the cross-age suffix scan
// (mzr.jl:172-181) over shared-memory copies and writes [-logL, G] (Nj + 3 numbers) to device memory and the pinned buffer.
// (Versions with the chain rule as a tail of the flat kernel -- per-age warps, then a product table summed by one block -- spent
// 5 us on 9600 uncoalesced 8-byte loads from one SM; profiles/r2_experiments.md section 4.)
__global__ void __launch_bounds__(kFinalizeThreads) sfh_finalize_hier_kernel(const FinalizeParams p) {
    __shared__ double part[kFinalizeThreads / 32][32], part2[kFinalizeThreads / 32][32];
    __shared__ double xch[kMaxExchangeRanks][32], xch2[kMaxExchangeRanks][32];   // the ranks' packets of one pass, [rank][member]
    __shared__ double sh[kFinalizeThreads / 32];
    __shared__ bool last;
    if (p.pdl_early) griddep_launch_dependents();
    griddep_wait();
    const HierTail &h = p.hier;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = kFinalizeThreads / 32;
    const int j = blockIdx.x;
    const unsigned long long epoch = p.peers ? __ldcg(p.epoch_ptr) + (p.epoch_from_fused ? 0ull : 1ull) : 0ull;
    const uint32_t ep32 = (uint32_t)epoch;
    const int64_t par = (int64_t)(epoch & 1ull);
    if (blockIdx.x == 0) {   // the shard's logL exists already (per-cluster partials of the stream kernel): finish / push it now
        const double all0 = early_logl(p, sh);
        if (warp == 0) {
            if (p.peers) push_logl(p, all0, lane, par, ep32);
            else if (lane == 0) p.out[0] = all0;
        }
    }
    if (p.want_G) {
        const int g0 = h.gptr[j], g1 = h.gptr[j + 1];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        // two members per lane and pass (g, g + 32): a group of up to 64 templates costs ONE round of dependent loads
#pragma unroll 1
        for (int base = g0; base < g1; base += 64) {
            const int gA_ = base + lane, gB_ = base + 32 + lane;
            const bool vA = gA_ < g1, vB = gB_ < g1;
            const int tA = vA ? h.gmem[gA_] : 0, tB = vB ? h.gmem[gB_] : 0;
            double wA[4] = {0.0, 0.0, 0.0, 0.0}, wB[4] = {0.0, 0.0, 0.0, 0.0};
            if (warp == 0) {   // (independent of G: requested before the partials)
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    if (vA) wA[f] = h.W[(int64_t)f * h.nt + gA_];
                    if (vB) wB[f] = h.W[(int64_t)f * h.nt + gB_];
                }
            }
            double sA = 0.0, sB = 0.0;
#pragma unroll 5
            for (int cl = warp; cl < p.n_clusters; cl += nw) {
                const double *row = p.gpart + (int64_t)cl * p.gstride;
                if (vA) sA += __ldcg(row + tA);
                if (vB) sB += __ldcg(row + tB);
            }
            part[warp][lane] = sA;
            part2[warp][lane] = sB;
            __syncthreads();
            double GA = 0.0, GB = 0.0;
            if (warp == 0) {
#pragma unroll
                for (int w = 0; w < nw; ++w) { GA += part[w][lane]; GB += part2[w][lane]; }
                if (p.peers) {
                    const int64_t slot = (par * p.nranks + p.rank) * p.vlen + 1;
#pragma unroll 1
                    for (int r = 0; r < p.nranks; ++r) {
                        if (vA) st_packet(reinterpret_cast<uint4 *>(p.peers[r]) + slot + tA, GA, ep32);
                        if (vB) st_packet(reinterpret_cast<uint4 *>(p.peers[r]) + slot + tB, GB, ep32);
                    }
                }
            }
            if (p.peers) {
                // every warp polls one rank's packets (lane = member), so the nranks waits run side by side instead of one
                // after another in warp 0 (8 dependent L2 round trips at 8 GPUs); warp 0 then adds them in rank order
                const uint4 *inbox = reinterpret_cast<const uint4 *>(p.peers[p.rank]) + par * p.nranks * p.vlen + 1;
#pragma unroll 1
                for (int r = warp; r < p.nranks; r += nw) {
                    xch[r][lane] = vA ? ld_packet_wait(inbox + (int64_t)r * p.vlen + tA, ep32) : 0.0;
                    xch2[r][lane] = vB ? ld_packet_wait(inbox + (int64_t)r * p.vlen + tB, ep32) : 0.0;
                }
                __syncthreads();
                if (warp == 0) {
                    GA = 0.0; GB = 0.0;
#pragma unroll 1
                    for (int r = 0; r < p.nranks; ++r) { GA += xch[r][lane]; GB += xch2[r][lane]; }   // rank order
                }
            }
            if (warp == 0) {
                // fullG_t = d logL / d r_t = -(M'r)_t
                if (vA) { p.out[1 + tA] = GA; a0 += (-GA) * wA[0]; a1 += (-GA) * wA[1]; a2 += (-GA) * wA[2]; a3 += (-GA) * wA[3]; }
                if (vB) { p.out[1 + tB] = GB; a0 += (-GB) * wB[0]; a1 += (-GB) * wB[1]; a2 += (-GB) * wB[2]; a3 += (-GB) * wB[3]; }
            }
            __syncthreads();
        }
        if (warp == 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
            if (lane == 0) { h.sums[j] = a0; h.sums[h.nj + j] = a1; h.sums[2 * h.nj + j] = a2; h.sums[3 * h.nj + j] = a3; }
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (p.dbg && tid == 0) p.dbg[1] = clock64();

    // ---------------- last block: logL, parameter sums, suffix scan ----------------
    __shared__ double s_dr[kHierTailAges], s_G[kHierTailAges], s_p[kHierTailAges], s_s[kHierTailAges], s_gA[kHierTailAges], s_gB[kHierTailAges];
    __shared__ int s_sidx[kHierTailAges];
    __shared__ double s_par[3];
    const int nj = h.nj;   // (<= kHierTailAges <= kFinalizeThreads)
    if (p.want_G) {
#pragma unroll 1
        for (int i = tid; i < nj; i += kFinalizeThreads) {
            s_sidx[i] = h.sidx[i];
            s_dr[i] = __ldcg(h.sums + i);                  // mzr.jl:166-167
            s_G[i] = -__ldcg(h.sums + nj + i);             // mzr.jl:188-190 / amr.jl:141
            s_p[i] = -__ldcg(h.sums + 2 * nj + i);         // mzr.jl:194-195
            s_s[i] = __ldcg(h.sums + 3 * nj + i);          // mzr.jl:206-207
            s_gA[i] = h.gA[i];
            s_gB[i] = h.gB[i];
        }
    }
    double all = 0.0;
    if (warp == 0) all = p.peers ? poll_logl(p, lane, par, ep32) : __ldcg(p.out);   // block 0 produced / pushed it at kernel start
    __syncthreads();   // the shared copies above are complete
    if (tid == 0) {
        p.out[0] = all;
        const double v = (all != 0.0) ? -all : __longlong_as_double(0x7ff0000000000000LL);  // fitting_base.jl:95, solvers.jl:31
        h.out[0] = v;
        if (h.out_host) h.out_host[0] = v;
        if (p.pkt_host) st_packet(reinterpret_cast<uint4 *>(p.pkt_host), v, (uint32_t)__ldcg(p.pkt_epoch));
        *p.ticket = 0u;
        if (p.peers && !p.epoch_from_fused) *p.epoch_ptr = epoch;
    }
    if (p.dbg && tid == 0) p.dbg[2] = clock64();
    if (!p.want_G) return;
    // parameter gradients (mzr.jl:196-198,201-208): warp 0, lane-strided partial sums + one xor tree each (fixed order)
    if (warp == 0) {
        double pa = 0.0, pb = 0.0, ps = 0.0;
#pragma unroll 1
        for (int i = lane; i < nj; i += 32) { pa += s_p[i] * s_gA[i]; pb += s_p[i] * s_gB[i]; ps -= s_s[i]; }
        pa = warp_sum(pa); pb = warp_sum(pb); ps = warp_sum(ps);
        if (lane == 0) { s_par[0] = pa; s_par[1] = pb; s_par[2] = ps; }
    }
    if (p.dbg && tid == 0) p.dbg[3] = clock64();
    if (h.kind == MH_POWERLAW_MZR) {
        // cum = reverse(cumsum(reverse(ksum[s])))  mzr.jl:172 ;  G[s[i-1]] -= cum[i]  :179-181: an inclusive SUFFIX sum over the
        // ages in sorted order, done as a log-step scan in shared memory (8 steps for 256 ages; the serial loop was 1.5 us of 60
        // dependent adds).  Pairs are added in a different order than a running sum: deterministic, and 1e-16 relative.
        __syncthreads();   // warp 0 is done with s_p / s_s
        double v = 0.0;
        if (tid < nj) v = s_dr[s_sidx[tid]];
        __syncthreads();
        if (tid < nj) s_p[tid] = v;
        __syncthreads();
#pragma unroll 1
        for (int d = 1; d < nj; d <<= 1) {
            const double add = (tid + d < nj) ? s_p[tid + d] : 0.0;
            __syncthreads();
            if (tid < nj) s_p[tid] += add;
            __syncthreads();
        }
        // s_p[i] = sum_{i' >= i} ksum[s[i']]:  G[s[i-1]] -= s_p[i]
        if (tid >= 1 && tid < nj) s_G[s_sidx[tid - 1]] -= s_p[tid];
    }
    __syncthreads();
    if (p.dbg && tid == 0) p.dbg[4] = clock64();
#pragma unroll 1
    for (int i = tid; i < nj + 3; i += kFinalizeThreads) {
        const double v = (i < nj) ? s_G[i] : (h.free_mask[i - nj] ? s_par[i - nj] : 0.0);   // fixed parameters receive 0 (mzr.jl:196,201)
        h.out[1 + i] = v;
        if (h.out_host) h.out_host[1 + i] = v;
        if (p.pkt_host) st_packet(reinterpret_cast<uint4 *>(p.pkt_host) + 1 + i, v, (uint32_t)__ldcg(p.pkt_epoch));
    }
}

// Upload of one evaluation's inputs WITHOUT a copy node: n doubles from the caller's mapped pinned buffer to device memory, plus
// the 64-bit epoch the host stored behind them (see FinalizeParams::pkt_host).  The dependent launch is released first, so the
// fused kernel becomes resident and its producers fill their rings while the 19 KB cross PCIe; its A warps wait for this grid.
// The device copy is written after griddepcontrol.wait only (a previous evaluation on the stream may still be reading it).
constexpr int kCopyInThreads = 256;
__global__ void __launch_bounds__(kCopyInThreads) sfh_copy_in_kernel(double *dst, const double *src_host, int64_t n, unsigned long long *epoch_dst) {
    griddep_launch_dependents();
    const int64_t i = ((int64_t)blockIdx.x * kCopyInThreads + threadIdx.x) * 2;
    double v0 = 0.0, v1 = 0.0;
    unsigned long long ep = 0ull;
    if (i + 1 < n) { const double2 v = __ldcv(reinterpret_cast<const double2 *>(src_host + i)); v0 = v.x; v1 = v.y; }
    else if (i < n) v0 = __ldcv(src_host + i);
    if (i == 0 && epoch_dst) ep = __ldcv(reinterpret_cast<const unsigned long long *>(src_host + n));
    griddep_wait();
    if (i < n) dst[i] = v0;
    if (i + 1 < n) dst[i + 1] = v1;
    if (i == 0 && epoch_dst) *epoch_dst = ep;
}

// ------------------------------------------------------------------------------------------
// two-pass kernels (unfused API + fallback when the fused tiling cannot hold T)
// ------------------------------------------------------------------------------------------
// composite!  C = M * coeffs   (fitting_base.jl:55-65): block = 128 bins x 4 template slices
template <typename S>
__global__ void __launch_bounds__(512) sfh_composite_kernel(const S *__restrict__ M, const StackLayout lay, int64_t nb, int64_t nt,
                                                            const double *__restrict__ coeffs,
                                                            double *__restrict__ out) {
    __shared__ double sh[4][128];
    const int bx = threadIdx.x & 127, sy = threadIdx.x >> 7;
    const int64_t i = (int64_t)blockIdx.x * 128 + bx;
    double acc = 0.0;
    if (i < nb) {
        const int64_t per = (nt + 3) / 4;
        const int64_t j0 = sy * per, j1 = (j0 + per < nt) ? j0 + per : nt;
        const int64_t cstride = lay.off(i, 1) - lay.off(i, 0);  // template-to-template stride (ld, or bt for panels)
        const S *col = M + lay.off(i, j0);
        int64_t j = j0;
        for (; j + 7 < j1; j += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (double)col[(int64_t)u * cstride];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma(v[u], coeffs[j + u], acc);
            col += 8 * cstride;
        }
        for (; j < j1; ++j, col += cstride) acc = fma((double)*col, coeffs[j], acc);
    }
    sh[sy][bx] = acc;
    __syncthreads();
    if (sy == 0 && i < nb) out[i] = (sh[0][bx] + sh[1][bx]) + (sh[2][bx] + sh[3][bx]);
}

__global__ void sfh_fill_kernel(double *x, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

// residual in place: C <- 1 - n/max(C,eps)   (fitting_base.jl:274-280)
__global__ void sfh_residual_kernel(double *__restrict__ C, const double *__restrict__ data, int64_t nb, double eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) {
        const double m = C[i];
        const double mc = (m < eps) ? eps : m;
        C[i] = 1.0 - data[i] / mc;
    }
}

// G_j = sign * sum_i M_ij r_i   (gemv 'T', fitting_base.jl:283): one warp per template
template <typename S>
__global__ void __launch_bounds__(256) sfh_gemvt_kernel(const S *__restrict__ M, const StackLayout lay, int64_t nb, int64_t nt,
                                                        const double *__restrict__ r, double sign,
                                                        double *__restrict__ G) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= nt) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int64_t i = lane;
    for (; i + 96 < nb; i += 128) {
        const double v0 = (double)M[lay.off(i, j)], v1 = (double)M[lay.off(i + 32, j)], v2 = (double)M[lay.off(i + 64, j)],
                     v3 = (double)M[lay.off(i + 96, j)];
        a0 = fma(v0, r[i], a0);
        a1 = fma(v1, r[i + 32], a1);
        a2 = fma(v2, r[i + 64], a2);
        a3 = fma(v3, r[i + 96], a3);
    }
    for (; i < nb; i += 32) a0 = fma((double)M[lay.off(i, j)], r[i], a0);
    const double s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) G[j] = sign * s;
}

// ------------------------------------------------------------------------------------------
// hierarchical prologue / epilogue (K5).  One block; one warp per age group, members visited in
// template order => deterministic.  Formulas cite src/fitting/hierarchical/{mzr,amr,dispersion_models}.jl
// ------------------------------------------------------------------------------------------
struct HierParams {
    int32_t kind;      // sfh_mh_kind
    int32_t nj;        // number of unique ages
    int64_t nt;
    double fixed[4];   // logMstar0 | T_max | T_max, solZ, Y_p, gamma
    uint8_t free_mask[4];
    const double *variables;  // [nj+3] device
    const double *logAge_u;   // [nj] unique ages, first-appearance order
    const double *MH;         // [nt]
    const int32_t *jidx;      // [nt] template -> age
    const int32_t *gptr;      // [nj+1] group offsets
    const int32_t *gmem;      // [nt] group members (template indices, ascending inside a group)
    const int32_t *sidx;      // [nj] sortperm(unique_logAge; rev=true)   mzr.jl:61
    // scratch (device)
    double *mu, *gA, *gB, *gM, *Asum, *cum;  // [nj]
    double *Ajk;                             // [nt]
    double *tmpj;                            // [4*nj]
    double *coeffs;                          // [nt] out of the prologue
    const double *fg_out;                    // [1+nt]: logL raw, +M'r  (input of the epilogue)
    double *out;                             // [1 + nj + 3]: -logL (guarded), G
    double *out_host;                        // nullable: mapped pinned host copy of `out`
    unsigned long long *pkt_epoch_out;       // nullable (folded path): device copy of the host's epoch word behind the variables
    int32_t pdl_early;                       // folded path: release the dependent launch at kernel start
    // batched launches (grid = chains): block k works on chain k, whose arrays sit k strides further on
    int64_t bs_vars, bs_scratch, bs_coeffs, bs_fg, bs_out;   // element strides; all 0 for a single evaluation
};

// chain blockIdx.x of a batched launch (identity for the single-evaluation launches: all strides are 0)
__device__ __forceinline__ HierParams hier_chain_params(const HierParams &p0) {
    HierParams p = p0;
    const int64_t k = blockIdx.x;
    p.variables += k * p0.bs_vars;
    p.mu += k * p0.bs_scratch; p.gA += k * p0.bs_scratch; p.gB += k * p0.bs_scratch; p.gM += k * p0.bs_scratch;
    p.Asum += k * p0.bs_scratch; p.cum += k * p0.bs_scratch; p.tmpj += k * p0.bs_scratch; p.Ajk += k * p0.bs_scratch;
    p.coeffs += k * p0.bs_coeffs;
    p.fg_out += k * p0.bs_fg;
    p.out += k * p0.bs_out;
    return p;
}

__device__ __forceinline__ double d_X_from_Z(double Z, double Yp, double gam) { return 1.0 - ((Yp + gam * Z) + Z); }

// mean metallicity + gradient: PowerLawMZR mzr.jl:275-280; LinearAMR amr.jl:206-209;
// LogarithmicAMR amr.jl:284-295 with MH_from_Z / dMH_dZ of src/utilities.jl:138-156
__device__ __forceinline__ void d_mh_eval(int kind, double alpha, double beta, const double *fx, double arg, double &mu,
                                          double &dA, double &dB, double &dM) {
    if (kind == MH_POWERLAW_MZR) {
        const double l = log10(arg) - fx[0];
        mu = beta + alpha * l;
        dA = l;
        dB = 1.0;
        dM = alpha / arg / 2.302585092994046;  // logten
    } else {
        const double age = exp10(arg - 9.0);
        const double dt = fx[0] - age;
        if (kind == MH_LINEAR_AMR) {
            mu = beta + alpha * dt;
            dA = dt;
            dB = 1.0;
            dM = 0.0;
        } else {
            const double Z = beta + alpha * dt;
            const double solZ = fx[1], Yp = fx[2], gam = fx[3];
            const double X = d_X_from_Z(Z, Yp, gam);
            mu = (X > 0.0) ? log10(Z / (X * solZ) * d_X_from_Z(solZ, Yp, gam)) : __longlong_as_double(0x7ff8000000000000LL);
            const double d = (Yp - 1.0) / (2.302585092994046 * Z * (Yp + Z + gam * Z - 1.0));
            dA = d * dt;
            dB = d;
            dM = 0.0;
        }
    }
}

constexpr int kHierThreads = 1024;
constexpr int kHierSmemAges = 1024;  // per-age scratch staged in shared memory up to this many unique ages

// calculate_coeffs: mzr.jl:50-79 / amr.jl:50-73
__global__ void __launch_bounds__(kHierThreads) sfh_hier_prologue_kernel(const HierParams p0) {
    const HierParams p = hier_chain_params(p0);
    griddep_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = kHierThreads / 32;
    const int nj = p.nj;
    const double alpha = p.variables[nj], beta = p.variables[nj + 1], sigma = p.variables[nj + 2];
    if (p.kind == MH_POWERLAW_MZR) {
        // cumsum(R[s])[invperm(s)]  mzr.jl:66 -- oldest first.  The scan is serial; its operands are staged in shared
        // memory first (a chain of dependent L2 round trips from one thread cost ~20 us at 60 ages)
        __shared__ double sR[kHierSmemAges];
        __shared__ int ssidx[kHierSmemAges];
        const bool staged = nj <= kHierSmemAges;
        if (staged) {
            for (int j = tid; j < nj; j += kHierThreads) { sR[j] = p.variables[j]; ssidx[j] = p.sidx[j]; }
            __syncthreads();
        }
        if (tid == 0) {
            double run = 0.0;
            for (int i = 0; i < nj; ++i) {
                const int j = staged ? ssidx[i] : p.sidx[i];
                run += staged ? sR[j] : p.variables[j];
                p.cum[j] = run;
            }
        }
        __syncthreads();
    }
    for (int j = tid; j < nj; j += kHierThreads) {
        const double arg = (p.kind == MH_POWERLAW_MZR) ? p.cum[j] : p.logAge_u[j];
        double mu, dA, dB, dM;
        d_mh_eval(p.kind, alpha, beta, p.fixed, arg, mu, dA, dB, dM);
        p.mu[j] = mu; p.gA[j] = dA; p.gB[j] = dB; p.gM[j] = dM;
    }
    __syncthreads();
    for (int j = warp; j < nj; j += nw) {
        const double mu = p.mu[j];
        const int g0 = p.gptr[j], g1 = p.gptr[j + 1];
        double a = 0.0;
        for (int g = g0 + lane; g < g1; g += 32) {
            const int t = p.gmem[g];
            const double z = (p.MH[t] - mu) / sigma;
            const double A = exp(-(z * z) / 2.0);  // dispersion_models.jl:92
            p.Ajk[t] = A;
            a += A;
        }
        a = warp_sum(a);
        if (lane == 0) p.Asum[j] = a;
        __syncwarp();
        const double Rj = p.variables[j];
        for (int g = g0 + lane; g < g1; g += 32) {
            const int t = p.gmem[g];
            p.coeffs[t] = p.Ajk[t] * Rj / a;  // mzr.jl:76
        }
    }
}

// calculate_coeffs (mzr.jl:50-79 / amr.jl:50-73) for ONE evaluation, spread over the grid: one warp per age group.  Besides the
// coefficients it prepares everything of the chain rule (mzr.jl:131-209) that does not depend on the gradient -- the factors
// W1..W4 that multiply fullG_jk in the four per-age sums, stored in AGE-GROUP order -- so that the finalize step is left with
// products and sums.  `vars_in` may be the caller's mapped pinned buffer: 63 doubles over PCIe instead of a memcpy node in front.
// Every global load is a DRAM round trip here (the stack stream has flushed L2) and every instruction a cold fetch (the kernel runs
// once per evaluation), so: the dependent load chain is gptr -> {MHg, gmem} (MHg = the metallicity grid in group order, built at
// sfh_hier_bind), the members' A_jk stay in shared memory between the passes, the cumulative mass runs over a sorted copy in
// shared memory, and each of the three member loops exists ONCE in the code (the first version, with the passes unrolled over a
// register window, was 3400 instructions and took 13 us; this one ~1000).
constexpr int kHierPro2Threads = 256;
constexpr int kHierPro2Keep = 128;   // members per group whose A_jk are kept in shared memory (beyond: recomputed)
__global__ void __launch_bounds__(kHierPro2Threads) sfh_hier_prologue2_kernel(const HierParams p, const double *vars_in, double *W,
                                                                               const double *MHg) {
    __shared__ double sv[kHierTailAges + 3], sRs[kHierTailAges];
    __shared__ int spos[kHierTailAges];
    __shared__ double sA[kHierPro2Threads / 32][kHierPro2Keep];
    if (p.pdl_early) griddep_launch_dependents();   // the fused kernel becomes resident and streams ahead; its A warps wait for this grid
    griddep_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = kHierPro2Threads / 32;
    const int nj = p.nj;
    const int j = blockIdx.x * nw + warp;
    // independent loads first: the group's extent and age
    int g0 = 0, g1 = 0;
    double age = 0.0;
    if (j < nj) { g0 = p.gptr[j]; g1 = p.gptr[j + 1]; age = p.logAge_u[j]; }
    // ... and the first 64 members' metallicities and template indices: their DRAM latency overlaps the PCIe read below
    const int gf = g0 + lane;
    const double mhp0 = (gf < g1) ? MHg[gf] : 0.0, mhp1 = (gf + 32 < g1) ? MHg[gf + 32] : 0.0;
    const int tp0 = (gf < g1) ? p.gmem[gf] : 0, tp1 = (gf + 32 < g1) ? p.gmem[gf + 32] : 0;
#pragma unroll 1
    for (int i = tid; i < nj + 3; i += kHierPro2Threads) {
        const double v = vars_in[i];
        sv[i] = v;
        if (blockIdx.x == 0) const_cast<double *>(p.variables)[i] = v;   // device copy for later kernels
    }
    // the host's epoch of this evaluation rides behind the variables (FinalizeParams::pkt_host)
    if (p.pkt_epoch_out && blockIdx.x == 0 && tid == 0) *p.pkt_epoch_out = __ldcv(reinterpret_cast<const unsigned long long *>(vars_in + nj + 3));
    int srt = -1;
    if (p.kind == MH_POWERLAW_MZR && tid < nj) srt = p.sidx[tid];          // (nj <= kHierTailAges <= block size)
    __syncthreads();
    if (srt >= 0) { sRs[tid] = sv[srt]; spos[srt] = tid; }                 // R in oldest-first order; where each age sits in it
    __syncthreads();
    if (j >= nj) return;
    const double alpha = sv[nj], beta = sv[nj + 1], sigma = sv[nj + 2];
    const double s2 = sigma * sigma, s3 = s2 * sigma;
    double arg = age;
    if (p.kind == MH_POWERLAW_MZR) {
        // cumsum(R[s])[invperm(s)]  mzr.jl:61-66 -- oldest first; this age's entry is the running sum up to its own position
        const int pos = spos[j];
        double run = 0.0;
#pragma unroll 4
        for (int i = 0; i <= pos; ++i) run += sRs[i];
        arg = run;
    }
    double mu, gA, gB, gM;
    d_mh_eval(p.kind, alpha, beta, p.fixed, arg, mu, gA, gB, gM);
    if (lane == 0) { p.mu[j] = mu; p.gA[j] = gA; p.gB[j] = gB; p.gM[j] = gM; }
    const double Rj = sv[j];
    double *myA = sA[warp];
    // pass 1: A_jk and their sum
    double a = 0.0;
#pragma unroll 1
    for (int g = g0 + lane; g < g1; g += 32) {
        const double mhv = (g == gf) ? mhp0 : (g == gf + 32 ? mhp1 : MHg[g]);
        const double z = (mhv - mu) / sigma;
        const double Av = exp(-(z * z) / 2.0);  // dispersion_models.jl:92
        if (g - g0 < kHierPro2Keep) myA[g - g0] = Av;
        a += Av;
    }
    const double Aj = warp_sum(a);
    if (lane == 0) p.Asum[j] = Aj;
    // pass 2 (per-age sums of the dispersion derivatives) and pass 3 (coefficients and chain-rule factors) share one loop body
    double kAR = 0.0, kmu = 0.0, ksg = 0.0;
    const double RA = Rj / Aj;
    const int64_t nt = p.nt;
#pragma unroll 1
    for (int pass = 2; pass <= 3; ++pass) {
#pragma unroll 1
        for (int g = g0 + lane; g < g1; g += 32) {
            const double d = ((g == gf) ? mhp0 : (g == gf + 32 ? mhp1 : MHg[g])) - mu;
            double Av;
            if (g - g0 < kHierPro2Keep) Av = myA[g - g0];
            else { const double z = d / sigma; Av = exp(-(z * z) / 2.0); }
            const double dAmu = Av * d / s2;      // dispersion_models.jl:99
            const double dAsg = Av * d * d / s3;  // dispersion_models.jl:98
            if (pass == 2) {
                kAR += dAmu * gM;                 // mzr.jl:162,164
                kmu += dAmu;
                ksg += dAsg;
            } else {
                const int t = (g == gf) ? tp0 : (g == gf + 32 ? tp1 : p.gmem[g]);
                const double coeff = Av * Rj / Aj;    // mzr.jl:76
                p.coeffs[t] = coeff;
                p.Ajk[t] = Av;
                if (p.kind == MH_POWERLAW_MZR) {
                    const double dAR = dAmu * gM;
                    W[g] = RA * (dAR - (Av * kAR / Aj));                            // mzr.jl:166-167
                    W[nt + g] = coeff / Rj + (dAR - (kAR * Av / Aj)) * Rj / Aj;     // mzr.jl:188-190
                } else {
                    W[g] = 0.0;
                    W[nt + g] = coeff / Rj;                                         // amr.jl:141
                }
                W[2 * nt + g] = RA * (dAmu - Av / Aj * kmu);                        // mzr.jl:194-195
                W[3 * nt + g] = RA * (dAsg - Av / Aj * ksg);                        // mzr.jl:206-207
            }
        }
        if (pass == 2) { kAR = warp_sum(kAR); kmu = warp_sum(kmu); ksg = warp_sum(ksg); }
    }
}

// chain rule: mzr.jl:124-210 / amr.jl:118-169.  fullG = d logL/d r = -(fg_out[1+t]).
__global__ void __launch_bounds__(kHierThreads) sfh_hier_epilogue_kernel(const HierParams p0, int want_G) {
    const HierParams p = hier_chain_params(p0);
    griddep_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = kHierThreads / 32;
    const int nj = p.nj;
    if (tid == 0) {
        const double logL = p.fg_out[0];
        const double v = (logL != 0.0) ? -logL : __longlong_as_double(0x7ff0000000000000LL);  // fitting_base.jl:95
        p.out[0] = v;
        if (p.out_host) p.out_host[0] = v;
    }
    if (!want_G) return;
    const double sigma = p.variables[nj + 2];
    const double s2 = sigma * sigma, s3 = s2 * sigma;
    double *ksumdr = p.tmpj, *same = p.tmpj + nj, *psum = p.tmpj + 2 * nj, *ssum = p.tmpj + 3 * nj;
    for (int j = warp; j < nj; j += nw) {
        const double mu = p.mu[j], Aj = p.Asum[j], Rj = p.variables[j], gM = p.gM[j];
        const int g0 = p.gptr[j], g1 = p.gptr[j + 1];
        double kAR = 0.0, kmu = 0.0, ksg = 0.0;
        for (int g = g0 + lane; g < g1; g += 32) {
            const int t = p.gmem[g];
            const double A = p.Ajk[t], d = p.MH[t] - mu;
            const double dAmu = A * d / s2;      // dispersion_models.jl:99
            const double dAsg = A * d * d / s3;  // dispersion_models.jl:98
            kAR += dAmu * gM;                    // mzr.jl:162,164
            kmu += dAmu;
            ksg += dAsg;
        }
        kAR = warp_sum(kAR); kmu = warp_sum(kmu); ksg = warp_sum(ksg);
        double a_dr = 0.0, a_same = 0.0, a_p = 0.0, a_s = 0.0;
        for (int g = g0 + lane; g < g1; g += 32) {
            const int t = p.gmem[g];
            const double A = p.Ajk[t], d = p.MH[t] - mu;
            const double dAmu = A * d / s2, dAsg = A * d * d / s3;
            const double fullG = -p.fg_out[1 + t];
            const double RA = Rj / Aj;
            if (p.kind == MH_POWERLAW_MZR) {
                const double dAR = dAmu * gM;
                a_dr += fullG * (RA * (dAR - (A * kAR / Aj)));                           // mzr.jl:166-167
                a_same += fullG * (p.coeffs[t] / Rj + (dAR - (kAR * A / Aj)) * Rj / Aj);  // mzr.jl:188-190
            } else {
                a_same += fullG * p.coeffs[t] / Rj;                                      // amr.jl:141
            }
            a_p += fullG * RA * (dAmu - A / Aj * kmu);                                   // mzr.jl:194-195
            a_s += fullG * RA * (dAsg - A / Aj * ksg);                                   // mzr.jl:206-207
        }
        a_dr = warp_sum(a_dr); a_same = warp_sum(a_same); a_p = warp_sum(a_p); a_s = warp_sum(a_s);
        if (lane == 0) { ksumdr[j] = a_dr; same[j] = a_same; psum[j] = -a_p; ssum[j] = a_s; }
    }
    __syncthreads();
    double *G = p.out + 1;
    // serial tails (suffix scan over ages, parameter sums) run on shared-memory copies: see the prologue's note
    __shared__ double s_dr[kHierSmemAges], s_G[kHierSmemAges];
    __shared__ int s_sidx[kHierSmemAges];
    __shared__ double s_par[3];
    const bool staged = nj <= kHierSmemAges;
    if (staged) {
        for (int j = tid; j < nj; j += kHierThreads) { s_dr[j] = ksumdr[j]; s_G[j] = -same[j]; s_sidx[j] = p.sidx[j]; }
    } else {
        for (int j = tid; j < nj; j += kHierThreads) G[j] = -same[j];
    }
    // parameter gradients: fixed-order (strided + xor tree) block sums  (mzr.jl:196-198,201-208)
    {
        double ga = 0.0, gb = 0.0, gs = 0.0;
        for (int j = tid; j < nj; j += kHierThreads) {
            ga += psum[j] * p.gA[j];
            gb += psum[j] * p.gB[j];
            gs -= ssum[j];
        }
        __shared__ double shp[kHierThreads / 32];
        const double ta = block_sum<kHierThreads>(ga, shp);
        const double tb = block_sum<kHierThreads>(gb, shp);
        const double ts = block_sum<kHierThreads>(gs, shp);
        if (tid == 0) { s_par[0] = ta; s_par[1] = tb; s_par[2] = ts; }
    }
    __syncthreads();
    if (tid == 0) {
        if (p.kind == MH_POWERLAW_MZR) {
            // cum = reverse(cumsum(reverse(ksum[s])))  mzr.jl:172 ;  G[s[i-1]] -= cum[i]  :179-181
            double run = 0.0;
            for (int i = nj - 1; i >= 1; --i) {
                if (staged) { run += s_dr[s_sidx[i]]; s_G[s_sidx[i - 1]] -= run; }
                else { run += ksumdr[p.sidx[i]]; G[p.sidx[i - 1]] -= run; }
            }
        }
        G[nj] = p.free_mask[0] ? s_par[0] : 0.0;
        G[nj + 1] = p.free_mask[1] ? s_par[1] : 0.0;
        G[nj + 2] = p.free_mask[2] ? s_par[2] : 0.0;
    }
    __syncthreads();
    if (staged) {
        for (int j = tid; j < nj; j += kHierThreads) G[j] = s_G[j];
    }
    if (p.out_host) {  // G is complete only after thread 0's serial tail above
        __syncthreads();
        for (int j = tid; j < nj + 3; j += kHierThreads) p.out_host[1 + j] = G[j];
    }
}

// ------------------------------------------------------------------------------------------
// dense BFGS inverse Hessian resident on the device (sfh_api.cu: DeviceHessian; opt-in, sfh_bfgs_opts.device_hessian):
// at 2400 templates the matrix is 46 MB -- one read for q = H g, one read + write for the rank-two update.
// ------------------------------------------------------------------------------------------
__global__ void sfh_bfgs_identity_kernel(double *H, int64_t n) {
    const int64_t tot = n * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x)
        H[e] = (e / n == e % n) ? 1.0 : 0.0;
}
// q_j = sum_i H[i + j n] g_i: one warp per column (H is symmetric, so column j is row j), lanes stride the column (coalesced),
// fixed-order shuffle tree => deterministic
__global__ void sfh_bfgs_symv_kernel(const double *__restrict__ H, const double *__restrict__ g, double *__restrict__ q, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += warps) {
        const double *col = H + j * n;
        double acc = 0.0;
        for (int64_t i = lane; i < n; i += 32) acc = fma(col[i], g[i], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) q[j] = acc;
    }
}
// H[i + j n] += (cs s_j - rho Hy_j) s_i - rho s_j Hy_i   (the BFGS update written as in csrc/sfh_drivers.h)
__global__ void sfh_bfgs_rank2_kernel(double *__restrict__ H, const double *__restrict__ s, const double *__restrict__ Hy, double rho,
                                      double cs, int64_t n) {
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
        const double a = cs * s[j] - rho * Hy[j], b = -rho * s[j];
        double *col = H + j * n;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
            col[i] += a * s[i] + b * Hy[i];
    }
}

// ------------------------------------------------------------------------------------------
// synthetic data on device: Philox4x32-10, counter = global linear element index
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double philox_u01(uint64_t idx, uint64_t seed, uint32_t stream) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), stream, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint64_t bits = (((uint64_t)r.x << 32) | r.y) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);  // [0,1)
}

template <typename S>
__global__ void sfh_fill_uniform_kernel(S *M, const StackLayout lay, int64_t rows, int64_t nt, int64_t row_begin,
                                        int64_t nbins_total, uint64_t seed, double scale) {
    const int64_t n = rows * nt;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, j = e / rows;
        const uint64_t gidx = (uint64_t)(row_begin + i) + (uint64_t)nbins_total * (uint64_t)j;
        M[lay.off(i, j)] = (S)(scale * philox_u01(gidx, seed, 0u));
    }
}

// n_i ~ Poisson(lambda_i): inversion for small lambda, PTRS (Hoermann 1993) otherwise
__global__ void sfh_poisson_kernel(const double *lam, double *out, int64_t rows, int64_t row_begin, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const double l = lam[i];
    const uint64_t gi = (uint64_t)(row_begin + i);
    uint32_t draw = 0;
    double k;
    if (!(l > 0.0)) {
        k = 0.0;
    } else if (l < 10.0) {
        const double L = exp(-l);
        double prod = philox_u01(gi, seed, 1u + draw++);
        k = 0.0;
        while (prod > L && k < 1000.0) {
            prod *= philox_u01(gi, seed, 1u + draw++);
            k += 1.0;
        }
    } else {
        const double slam = sqrt(l), loglam = log(l);
        const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
        const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
        k = floor(l);
        for (int tries = 0; tries < 256; ++tries) {
            const double U = philox_u01(gi, seed, 1u + draw++) - 0.5;
            const double V = philox_u01(gi, seed, 1u + draw++);
            const double us = 0.5 - fabs(U);
            const double kk = floor((2.0 * a / us + b) * U + l + 0.43);
            if (us >= 0.07 && V <= vr) { k = kk; break; }
            if (kk < 0.0 || (us < 0.013 && V > us)) continue;
            if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -l + kk * loglam - lgamma(kk + 1.0)) { k = kk; break; }
        }
    }
    out[i] = k;
}

// column-major block (rows x ncols, leading dimension rows, columns j0..j0+ncols) <-> the stack's device layout
template <typename S>
__global__ void sfh_relayout_kernel(S *stack, const StackLayout lay, S *colblk, int64_t rows, int64_t j0, int64_t ncols,
                                    int to_stack) {
    const int64_t n = rows * ncols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, jj = e / rows;
        if (to_stack) stack[lay.off(i, j0 + jj)] = colblk[e];
        else colblk[e] = stack[lay.off(i, j0 + jj)];
    }
}

// data conversion on upload (Int64 / Float32 -> double)
template <typename D>
__global__ void sfh_to_double_kernel(const D *in, double *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

// Float32 stacks: does any element have its sign bit set, or is any Inf / NaN?  (decides whether the stream kernel may use its
// conversion-free unpack, sfh_fused2.cuh)
__global__ void sfh_check_f32_kernel(const uint32_t *bits, int64_t n, int *flag) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = bits[i];
        bad |= (b & 0x80000000u) != 0u || (b & 0x7f800000u) == 0x7f800000u;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}

// writes >L2 bytes: used between timed evaluations (bench hygiene)
__global__ void sfh_l2_flush_kernel(float4 *buf, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace sfh
