// sfh_packets.h -- the HOST side of "completion by packets" (host-only C++, no CUDA: also built by tests/native_host_sanitize.cpp).
//
// A host-synchronous evaluation (sfh_eval_fg / sfh_eval_fg_hier; reference callers: solvers.jl:88-89, hmc_sample.jl:24-37) does not
// synchronise the stream.  The finalize kernel stores every result into pinned host memory as a 16-byte packet
//     { lo32(value), epoch, hi32(value), epoch }
// i.e. two 8-byte halves that each carry the epoch of the evaluation the host asked for (st_packet, csrc/sfh_small.cuh).  An 8-byte
// aligned half is written by a single PCIe write and read by a single host load, so a packet validates itself however the two halves
// are ordered or delayed on the way: it is complete exactly when BOTH halves show this evaluation's epoch, and a stale packet of an
// earlier evaluation can never be taken for it.  No fence, no flag, no copy.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace sfh_packets {

enum StreamState { kRunning = 0, kDrained = 1, kFailed = -1 };   // what the stream query of the caller reports
enum WaitResult { kOk = 0, kMissing = 1, kStreamError = 2 };

inline void encode(uint64_t *packet, double v, uint32_t epoch) {   // (tests and host-side emulation; the device has its own)
    uint64_t bits;
    memcpy(&bits, &v, 8);
    packet[0] = (bits & 0xffffffffull) | ((uint64_t)epoch << 32);
    packet[1] = (bits >> 32) | ((uint64_t)epoch << 32);
}

// One packet: true (and *v) when both halves carry `epoch`.
inline bool try_read(const uint64_t *packet, uint32_t epoch, double *v) {
    const uint64_t a = __atomic_load_n(packet, __ATOMIC_ACQUIRE), b = __atomic_load_n(packet + 1, __ATOMIC_ACQUIRE);
    if ((uint32_t)(a >> 32) != epoch || (uint32_t)(b >> 32) != epoch) return false;
    const uint64_t bits = (a & 0xffffffffull) | (b << 32);
    memcpy(v, &bits, 8);
    return true;
}

// Wait for packet 0 (-> *first, nullable) and, when rest != nullptr, for the n packets from index rest_at on (-> rest[0..n)).
// `query` is called every `query_every` unsuccessful probes and tells whether the stream is still running: once it has drained, the
// packets must all be there on the next pass (kMissing otherwise: an evaluation that delivered nothing must fail, not hang); a
// stream error ends the wait at once.  *missing = index of the packet that was being waited for.
template <typename Query>
int wait(const uint64_t *packets, uint32_t epoch, double *first, double *rest, size_t n, size_t rest_at, Query &&query,
         size_t *missing, uint64_t query_every = 8192) {
    const size_t total = 1 + (rest ? n : 0);
    uint64_t spins = 0;
    bool drained = false;
    for (size_t j = 0; j < total;) {
        const size_t at = j == 0 ? 0 : rest_at + j - 1;
        double v;
        if (try_read(packets + 2 * at, epoch, &v)) {
            if (j == 0) { if (first) *first = v; } else rest[j - 1] = v;
            ++j;
            continue;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if (++spins % query_every == 0) {
            if (missing) *missing = at;
            if (drained) return kMissing;
            const int st = query();
            if (st == kDrained) drained = true;
            else if (st == kFailed) return kStreamError;
        }
    }
    return kOk;
}

}  // namespace sfh_packets
