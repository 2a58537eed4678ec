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

struct Cfg {
    int N;        // filters
    int shift;    // A start shift in rows of 16 B
    int pattern;  // 0: one MMA kind (N) repeated;  1: dual (N=2n, then N=n);  2: triple (n, n, n) = the plain 3-term split
    int mb;       // number of D tiles rotated through (1 = same accumulator every time)
    int reps;
    int a_pitch_rows;  // plane size in rows (LBO = a_pitch_rows * 16)
    int branch_free;   // 1: no elect branch, predicated MMA
};

__global__ void __launch_bounds__(128) probe(Cfg cfg, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // finite bf16 values
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    if (warp == 0) {
        const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lbo = (uint32_t)cfg.a_pitch_rows << 16;
        const uint32_t a0 = ((smem_u32(smem) >> 4) + (uint32_t)cfg.shift) | a_lbo;          // A planes: [chunk0][chunk1][lo chunk0][lo chunk1]
        const uint32_t a_lo_off = 2u * (uint32_t)cfg.a_pitch_rows;
        const uint32_t b_base = (smem_u32(smem + 128 * 1024) >> 4);
        const int n = cfg.N;
        const uint32_t b_lbo = (uint32_t)(2 * n) << 16;
        const uint32_t b0 = b_base | b_lbo;
        const uint32_t id1 = make_idesc(n), id2 = make_idesc(2 * n);
        const uint32_t dstep = (uint32_t)(cfg.pattern == 1 ? 2 * n : n);
        long long t0 = clock64();
        if (cfg.branch_free == 2) {  // ONE elect for the whole loop nest
            if (elect_one()) {
                for (int rep = 0; rep < cfg.reps; ++rep) {
                    uint32_t a = a0, d = tmem;
                    for (int mb = 0; mb < cfg.mb; ++mb, a += 128u, d += dstep) {
                        if (cfg.pattern == 0) {
                            mma(d, make_desc(a, hi), make_desc(b0, hi), id1, 1u);
                        } else if (cfg.pattern == 1) {
                            mma(d, make_desc(a, hi), make_desc(b0, hi), id2, 1u);
                            mma(d, make_desc(a + a_lo_off, hi), make_desc(b0, hi), id1, 1u);
                        } else {
                            mma(d, make_desc(a, hi), make_desc(b0, hi), id1, 1u);
                            mma(d, make_desc(a + a_lo_off, hi), make_desc(b0, hi), id1, 1u);
                            mma(d, make_desc(a, hi), make_desc(b0 + (uint32_t)n, hi), id1, 1u);
                        }
                    }
                }
            }
            __syncwarp();
        } else if (cfg.branch_free) {
            for (int rep = 0; rep < cfg.reps; ++rep) {
                uint32_t a = a0, d = tmem;
                for (int mb = 0; mb < cfg.mb; ++mb, a += 128u, d += dstep) {
                    if (cfg.pattern == 0) {
                        mma_e(d, make_desc(a, hi), make_desc(b0, hi), id1, 1u);
                    } else if (cfg.pattern == 1) {
                        mma_e(d, make_desc(a, hi), make_desc(b0, hi), id2, 1u);
                        mma_e(d, make_desc(a + a_lo_off, hi), make_desc(b0, hi), id1, 1u);
                    } else {
                        mma_e(d, make_desc(a, hi), make_desc(b0, hi), id1, 1u);
                        mma_e(d, make_desc(a + a_lo_off, hi), make_desc(b0, hi), id1, 1u);
                        mma_e(d, make_desc(a, hi), make_desc(b0 + (uint32_t)n, hi), id1, 1u);
                    }
                }
            }
        } else
        for (int rep = 0; rep < cfg.reps; ++rep) {
            if (elect_one()) {
                uint32_t a = a0, d = tmem;
                for (int mb = 0; mb < cfg.mb; ++mb, a += 128u, d += dstep) {
                    if (cfg.pattern == 0) {
                        mma(d, make_desc(a, hi), make_desc(b0, hi), id1, 1u);
                    } else if (cfg.pattern == 1) {
                        mma(d, make_desc(a, hi), make_desc(b0, hi), id2, 1u);
                        mma(d, make_desc(a + a_lo_off, hi), make_desc(b0, hi), id1, 1u);
                    } else {
                        mma(d, make_desc(a, hi), make_desc(b0, hi), id1, 1u);
                        mma(d, make_desc(a + a_lo_off, hi), make_desc(b0, hi), id1, 1u);
                        mma(d, make_desc(a, hi), make_desc(b0 + (uint32_t)n, hi), id1, 1u);
                    }
                }
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        // wait for completion
        uint32_t ok = 0;
        for (int i = 0; i < (1 << 24) && !ok; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t1 = clock64();
        if (tid == 0) *cycles = ok ? t1 - t0 : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}

static void run(const char* what, Cfg c) {
    long long* dC;
    CK(cudaMalloc(&dC, 8));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe<<<1, 128, 200 * 1024>>>(c, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", what, cudaGetErrorString(e)); exit(3); }
    long long cyc;
    CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
    const int per = c.pattern == 0 ? 1 : (c.pattern == 1 ? 2 : 3);
    const double n_mma = (double)c.reps * c.mb * per;
    printf("%-44s N=%3d shift=%3d mb=%d pitch=%4d : %8lld clk, %6.1f clk/MMA, %7.1f clk per (A-tile x all terms)\n", what, c.N, c.shift, c.mb, c.a_pitch_rows, cyc,
           cyc / n_mma, cyc / ((double)c.reps * c.mb));
    cudaFree(dC);
}

int main() {
    const int reps = 200;
    for (int N : {16, 32, 48, 64, 96, 128, 256}) run("ONE-ELECT single kind, one D", {N, 0, 0, 1, reps, 640, 2});
    for (int N : {16, 64, 128}) run("ONE-ELECT single kind, 4 rotating D", {N, 0, 0, 4, reps, 640, 2});
    for (int sh : {1, 8}) run("ONE-ELECT single N=16, shifted A, 4 D", {16, sh, 0, 4, reps, 640, 2});
    for (int N : {16, 32, 48, 64, 128}) run("ONE-ELECT dual (2n then n)", {N, 0, 1, N <= 64 ? 4 : 2, reps, 640, 2});
    for (int N : {16, 48, 64, 128, 256}) run("ONE-ELECT triple (n, n, n)", {N, 0, 2, N <= 128 ? 4 : 2, reps, 640, 2});
    for (int N : {48, 128}) run("ONE-ELECT triple, shift 1", {N, 1, 2, N <= 128 ? 4 : 2, reps, 640, 2});
    for (int N : {16, 32, 48, 64, 96, 128, 256}) run("BRANCH-FREE single kind, one D", {N, 0, 0, 1, reps, 640, 1});
    for (int N : {16, 64, 128}) run("BRANCH-FREE single kind, 4 rotating D", {N, 0, 0, 4, reps, 640, 1});
    for (int sh : {1, 7, 8, 126}) run("BRANCH-FREE single N=64, shifted A, 4 D", {64, sh, 0, 4, reps, 640, 1});
    for (int sh : {1, 8}) run("BRANCH-FREE single N=16, shifted A, 4 D", {16, sh, 0, 4, reps, 640, 1});
    for (int N : {16, 32, 48, 64, 128}) run("BRANCH-FREE dual (2n then n)", {N, 0, 1, N <= 64 ? 4 : 2, reps, 640, 1});
    for (int N : {16, 64}) run("BRANCH-FREE dual, shift 1", {N, 1, 1, 4, reps, 640, 1});
    for (int N : {16, 48, 64, 128, 256}) run("BRANCH-FREE triple (n, n, n)", {N, 0, 2, N <= 128 ? 4 : 2, reps, 640, 1});
    for (int N : {48, 128}) run("BRANCH-FREE triple, shift 1", {N, 1, 2, N <= 128 ? 4 : 2, reps, 640, 1});
    for (int N : {16, 32, 48, 64, 96, 128, 256}) run("single kind, aligned A, one D", {N, 0, 0, 1, reps, 640});
    for (int N : {16, 64, 128}) run("single kind, aligned A, 4 rotating D", {N, 0, 0, N <= 128 ? 4 : 2, reps, 640});
    for (int sh : {1, 2, 4, 7, 8, 9, 126, 128}) run("single kind N=64, shifted A, 4 D", {64, sh, 0, 4, reps, 640});
    for (int sh : {1, 8}) run("single kind N=16, shifted A, 4 D", {16, sh, 0, 4, reps, 640});
    for (int pitch : {640, 648, 768, 772}) run("single kind N=64, plane pitch (LBO) varied", {64, 0, 0, 4, reps, pitch});
    for (int N : {16, 32, 48, 64, 128}) run("dual (2n then n), aligned", {N, 0, 1, N <= 64 ? 4 : 2, reps, 640});
    for (int N : {16, 64}) run("dual (2n then n), shift 1", {N, 1, 1, 4, reps, 640});
    for (int N : {16, 48, 64, 128, 256}) run("triple (n, n, n), aligned", {N, 0, 2, N <= 128 ? 4 : 2, reps, 640});
    for (int N : {48, 128}) run("triple (n, n, n), shift 1", {N, 1, 2, N <= 128 ? 4 : 2, reps, 640});
    printf("RATE PROBE OK\n");
    return 0;
}
