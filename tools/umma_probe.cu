// umma_probe.cu -- hardware probe for the assumptions the tcgen05 conv engine is built on (run on a B200):
//   1. K-major, no-swizzle (INTERLEAVE) smem operand layout: rows 16 B apart inside an 8-row core matrix,
//      SBO between 8-row groups, LBO between the two 16-byte K chunks of one K=16 MMA;
//   2. the descriptor start address may be ANY multiple of 16 B, i.e. the A operand can be shifted by whole rows
//      (pixels) without moving data -- the im2col-free tap shift;
//   3. LBO is free: the two K chunks of one MMA may live in unrelated planes;
//   4. cycles per MMA as a function of N (is a small-N MMA bounded by the 128 B/clk shared-memory read of A?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run: timeout 60 ./umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}

__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 1u << 7;                    // a_format = BF16
    d |= 1u << 10;                   // b_format = BF16
    d |= (uint32_t)(N >> 3) << 17;   // n_dim
    d |= (uint32_t)(M >> 4) << 24;   // m_dim
    return d;                        // a_major = b_major = K (0)
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int max_spins) {
    for (int i = 0; i < max_spins; ++i) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

struct ProbeCfg {
    int N;          // MMA N
    int ksteps;     // K = 16 * ksteps
    int shift;      // A row shift (rows) applied through the descriptor start address
    int a_rows;     // rows physically stored per A plane (>= 128 + shift)
    int swap_lbo;   // 1: swap the LBO/SBO roles (hypothesis test; should FAIL)
    int split_planes;  // 1: odd K chunks are stored in a second, unrelated region (arbitrary LBO)
    int reps;       // timing: repeat the whole K loop this many times
};

// A_g: [a_rows][K] bf16 row-major, B_g: [N][K] bf16 row-major, D_g: [128][N] fp32
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* A_g, const __nv_bfloat16* B_g, float* D_g, ProbeCfg cfg,
                                                    long long* cycles, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = 16 * cfg.ksteps, nchunk = 2 * cfg.ksteps;
    const uint32_t a_plane = cfg.a_rows * 16, b_plane = cfg.N * 16;
    uint8_t* sA = smem;
    uint8_t* sA2 = sA + nchunk * a_plane + 4096;  // "unrelated" region for odd chunks when split_planes
    uint8_t* sB = sA2 + nchunk * a_plane + 1024;
    auto a_addr = [&](int chunk) -> uint8_t* {
        if (cfg.split_planes && (chunk & 1)) return sA2 + (chunk >> 1) * a_plane;
        if (cfg.split_planes) return sA + (chunk >> 1) * a_plane;
        return sA + chunk * a_plane;
    };
    // fill A: element (r, k) -> plane(k/8) + r*16 + (k%8)*2
    for (int i = tid; i < cfg.a_rows * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16*>(a_addr(k >> 3) + r * 16 + (k & 7) * 2) = A_g[i];
    }
    for (int i = tid; i < cfg.N * K; i += 128) {
        const int n = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 3) * b_plane + n * 16 + (k & 7) * 2) = B_g[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_bf16(128, cfg.N);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        // descriptors precomputed: the issue loop below is 4 MMAs + loop overhead (the first version rebuilt both
        // descriptors per MMA and measured 183 cyc/MMA for every N, i.e. single-thread issue overhead, not the tensor pipe)
        uint64_t ad[4], bd[4];
        for (int j = 0; j < 4; ++j) {
            const int jj = j % cfg.ksteps;
            uint32_t a0 = smem_u32(a_addr(2 * jj)) + cfg.shift * 16;
            uint32_t lbo = smem_u32(a_addr(2 * jj + 1)) - smem_u32(a_addr(2 * jj));
            ad[j] = cfg.swap_lbo ? make_desc(a0, 128, lbo) : make_desc(a0, lbo, 128);
            uint32_t b0 = smem_u32(sB + 2 * jj * b_plane);
            bd[j] = cfg.swap_lbo ? make_desc(b0, 128, b_plane) : make_desc(b0, b_plane, 128);
        }
        t0 = clock64();
        if (cfg.ksteps == 4) {
            mma_bf16(tmem, ad[0], bd[0], idesc, 0u);
            mma_bf16(tmem, ad[1], bd[1], idesc, 1u);
            mma_bf16(tmem, ad[2], bd[2], idesc, 1u);
            mma_bf16(tmem, ad[3], bd[3], idesc, 1u);
            for (int rep = 1; rep < cfg.reps; ++rep) {
                mma_bf16(tmem, ad[0], bd[0], idesc, 1u);
                mma_bf16(tmem, ad[1], bd[1], idesc, 1u);
                mma_bf16(tmem, ad[2], bd[2], idesc, 1u);
                mma_bf16(tmem, ad[3], bd[3], idesc, 1u);
            }
        } else {
            for (int rep = 0; rep < cfg.reps; ++rep)
                for (int j = 0; j < cfg.ksteps; ++j) mma_bf16(tmem, ad[j], bd[j], idesc, (rep | j) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const bool ok = mbar_wait(smem_u32(&bar), 0, 1 << 22);
    if (tid == 0) {
        t1 = clock64();
        *cycles = t1 - t0;
    }
    if (!ok) {
        if (tid == 0) *status = 1;  // timed out
    } else {
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int c0 = 0; c0 < cfg.N; c0 += 8) {
            uint32_t v[8];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int q = 0; q < 8; ++q) D_g[(size_t)tid * cfg.N + c0 + q] = __uint_as_float(v[q]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static bool run(const char* name, ProbeCfg cfg, bool expect_pass) {
    const int K = 16 * cfg.ksteps;
    std::vector<__nv_bfloat16> A((size_t)cfg.a_rows * K), B((size_t)cfg.N * K);
    std::vector<float> Af(A.size()), Bf(B.size());
    srand(123);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((rand() % 17) - 8); A[i] = __float2bfloat16(Af[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)((rand() % 13) - 6) * 0.25f; B[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB; float* dD; long long* dC; int* dS;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * cfg.N * 4));
    CK(cudaMalloc(&dC, 8)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * cfg.N * 4)); CK(cudaMemset(dS, 0, 4));
    const size_t smem = (size_t)2 * (2 * cfg.ksteps * cfg.a_rows * 16) + 4096 + 1024 + (size_t)2 * cfg.ksteps * cfg.N * 16 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, cfg, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(3); }
    std::vector<float> D(128 * cfg.N); long long cyc; int st;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < cfg.N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)(m + cfg.shift) * K + k] * Bf[(size_t)n * K + k];
            ref *= cfg.reps;
            double err = fabs(ref - D[m * cfg.N + n]);
            if (!(err <= maxerr)) maxerr = err;
        }
    const bool pass = (st == 0) && (maxerr == 0.0);
    printf("%-34s N=%3d K=%4d shift=%2d reps=%3d : %s (maxerr %.3g, status %d)  %lld cyc total, %.1f cyc/MMA  [%s]\n", name, cfg.N, K,
           cfg.shift, cfg.reps, pass ? "MATCH" : "MISMATCH", maxerr, st, cyc, (double)cyc / (cfg.reps * cfg.ksteps),
           pass == expect_pass ? "as expected" : "UNEXPECTED");
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC); cudaFree(dS);
    return pass == expect_pass;
}

int main() {
    bool ok = true;
    ok &= run("basic", {16, 1, 0, 160, 0, 0, 1}, true);
    ok &= run("basic K=64", {16, 4, 0, 160, 0, 0, 1}, true);
    ok &= run("N=48", {48, 2, 0, 160, 0, 0, 1}, true);
    ok &= run("N=96 K=48", {96, 3, 0, 160, 0, 0, 1}, true);
    ok &= run("N=256", {256, 2, 0, 160, 0, 0, 1}, true);
    ok &= run("row shift 1 (16 B)", {32, 2, 1, 160, 0, 0, 1}, true);
    ok &= run("row shift 3", {32, 2, 3, 160, 0, 0, 1}, true);
    ok &= run("row shift 21", {32, 2, 21, 160, 0, 0, 1}, true);
    ok &= run("split planes (free LBO)", {32, 2, 5, 160, 0, 1, 1}, true);
    // (swapping the LBO/SBO roles reads out of bounds -> "illegal memory access" on B200: confirms the field roles)
    // timing: cycles per MMA vs N (K=16 per MMA), 64 MMAs back to back
    for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
        char nm[64];
        snprintf(nm, sizeof(nm), "timing N=%d", N);
        ok &= run(nm, {N, 4, 0, 160, 0, 0, 8}, true);
        ok &= run(nm, {N, 4, 0, 160, 0, 0, 64}, true);
    }
    printf(ok ? "PROBE OK\n" : "PROBE FAILED\n");
    return ok ? 0 : 1;
}
