// bw_probe.cu -- what read-only HBM bandwidth can ANY kernel reach on this B200 for a 1.15 GB single pass?
// (1) grid-stride LDG.128 sum; (2) 1-D bulk-TMA (cp.async.bulk) into a shared-memory ring, consumer just touches it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) ldg_sum(const double2 *__restrict__ p, size_t n2, double *out) {
    double a = 0, b = 0, c = 0, d = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n2; i += 4 * st) {
        double2 v0 = p[i], v1 = p[i + st], v2 = p[i + 2 * st], v3 = p[i + 3 * st];
        a += v0.x + v0.y; b += v1.x + v1.y; c += v2.x + v2.y; d += v3.x + v3.y;
    }
    for (; i < n2; i += st) { double2 v = p[i]; a += v.x + v.y; }
    a += b + c + d;
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(~0u, a, o);
    if ((threadIdx.x & 31) == 0 && a == 12345.678) out[0] = a;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(288) bulk_ring(const char *__restrict__ p, size_t nchunks, double *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = (uint64_t *)(sm + (size_t)STAGES * CHUNK), *empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait = [](uint64_t *b, uint32_t ph) {
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
    };
    if (warp == 8) {
        if (lane == 0) {
            int slot = 0; uint32_t round = 0;
            for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
                if (round) wait(&empty[slot], (round - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[slot])), "r"(CHUNK) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)slot * CHUNK)),
                             "l"(p + c * CHUNK), "r"(CHUNK), "r"(s32(&full[slot])) : "memory");
                if (++slot == STAGES) { slot = 0; ++round; }
            }
        }
    } else {
        int slot = 0; uint32_t ph = 0; double acc = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            wait(&full[slot], ph);
            const double2 v = *(const double2 *)(sm + (size_t)slot * CHUNK + tid * 16);
            acc += v.x + v.y;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[slot])) : "memory");
            if (++slot == STAGES) { slot = 0; ph ^= 1; }
        }
        if (acc == 12345.678) out[0] = acc;
    }
}

int main() {
    const size_t bytes_list[2] = {1152000000ull, 8ull << 30};
    double *out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t bytes : bytes_list) {
        char *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
        float ms;
        for (int grid : {148 * 2, 148 * 4, 148 * 8}) {
            for (int r = 0; r < 3; ++r) ldg_sum<<<grid, 512>>>((const double2 *)buf, bytes / 16, out);
            cudaEventRecord(e0); for (int r = 0; r < 10; ++r) ldg_sum<<<grid, 512>>>((const double2 *)buf, bytes / 16, out); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("ldg_sum   bytes=%zu grid=%d  %.1f us  %.0f GB/s\n", bytes, grid, ms * 100, bytes / (ms / 10 * 1e-3) / 1e9);
        }
        {
            constexpr int ST = 26, CH = 4096;
            const size_t smem = (size_t)ST * CH + 2 * ST * 8;
            cudaFuncSetAttribute(bulk_ring<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int grid : {148, 296}) {
                for (int r = 0; r < 3; ++r) bulk_ring<ST, CH><<<grid, 288, smem>>>(buf, bytes / CH, out);
                cudaEventRecord(e0); for (int r = 0; r < 10; ++r) bulk_ring<ST, CH><<<grid, 288, smem>>>(buf, bytes / CH, out); cudaEventRecord(e1);
                cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                printf("bulk_ring 4KBx26 bytes=%zu grid=%d  %.1f us  %.0f GB/s\n", bytes, grid, ms * 100, bytes / (ms / 10 * 1e-3) / 1e9);
            }
        }
        {
            constexpr int ST = 12, CH = 16384;
            const size_t smem = (size_t)ST * CH + 2 * ST * 8;
            cudaFuncSetAttribute(bulk_ring<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int r = 0; r < 3; ++r) bulk_ring<ST, CH><<<148, 288, smem>>>(buf, bytes / CH, out);
            cudaEventRecord(e0); for (int r = 0; r < 10; ++r) bulk_ring<ST, CH><<<148, 288, smem>>>(buf, bytes / CH, out); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("bulk_ring 16KBx12 bytes=%zu grid=148  %.1f us  %.0f GB/s\n", bytes, ms * 100, bytes / (ms / 10 * 1e-3) / 1e9);
        }
        cudaFree(buf);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
