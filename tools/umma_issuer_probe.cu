// umma_rate_probe.cu -- how fast does tcgen05.mma (kind::f16, M=128, K=16, SS operands, K-major no-swizzle) really go on B200?
// The issue loop is warp-uniform with one elected lane (descriptors in uniform registers), unlike tools/umma_probe.cu whose
// divergent single-thread loop costs ~47 clk per MMA by itself.  Questions:
//   1. cycles per MMA vs N with an ALIGNED A start;          2. ... with the A start shifted by r rows (16 B each): the im2col-free tap
//   3. the "dual" pattern (N = 2n then N = n into the same D); 4. rotating D tiles (the M-block loop of the conv engine)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_probe umma_rate_probe.cu ; run: timeout 60 ./umma_rate_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
// branch-free form: every lane executes, the instruction itself is predicated on the elected lane
__device__ __forceinline__ void mma_e(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}


// two_issuer_probe: do two issuing threads (different warps, different accumulators) overlap their tcgen05.mma issue cost?
__global__ void __launch_bounds__(128) probe2(int N, int reps, int n_issuers, int per_rep, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    long long t0 = clock64();
    if (warp < n_issuers) {
        const uint32_t tmem = tmem_base_s + (uint32_t)(warp * 128);   // own accumulator columns
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lbo = 640u << 16;
        const uint32_t a0 = ((smem_u32(smem) >> 4) + (uint32_t)(warp * 16)) | a_lbo;
        const uint32_t b_lbo = (uint32_t)(2 * N) << 16;
        const uint32_t b0 = (smem_u32(smem + 128 * 1024) >> 4) | b_lbo;
        const uint32_t id1 = make_idesc(N);
        if (elect_one()) {
            for (int rep = 0; rep < reps; ++rep) {
                uint32_t d = tmem;
                for (int k = 0; k < per_rep; ++k) {
                    mma(d + (uint32_t)((k & 1) * 64), make_desc(a0 + (uint32_t)k, hi), make_desc(b0, hi), id1, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        for (int i = 0; i < (1 << 24) && !ok; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0) : "memory");
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}

int main() {
    long long* dC;
    CK(cudaMalloc(&dC, 8));
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int reps = 200, per_rep = 12;
    for (int N : {16, 32, 64, 96, 128, 256}) {
        for (int ni : {1, 2, 3}) {
            if (N == 256 && ni > 2) continue;
            probe2<<<1, 128, 200 * 1024>>>(N > 64 && ni > 1 ? N : N, reps, ni, per_rep, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("KERNEL ERROR %s\n", cudaGetErrorString(e)); return 3; }
            long long cyc;
            CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
            const double n_mma = (double)reps * per_rep * ni;
            printf("N=%3d issuers=%d : %8lld clk total, %6.1f clk per MMA overall (%6.1f per issuer-MMA)\n", N, ni, cyc, cyc / n_mma, cyc / ((double)reps * per_rep));
        }
    }
    printf("TWO ISSUER PROBE OK\n");
    return 0;
}
