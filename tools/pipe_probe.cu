// pipe_probe.cu -- per-SM issue rates of the instructions the fused S2D kernel is built from (run on a B200):
//   legacy tensor path  mma.sync m16n8k16 bf16 / m16n8k8 tf32 (SASS HMMA), ldmatrix, min/max (2- and 3-input, NaN variants),
//   FFMA and packed FFMA2.  Every loop keeps 8 independent dependency chains per warp; the grid fills every SM with
//   `warps` warps.  Output: warp-instructions per clock per SM (measured with clock64 inside the kernel, max over CTAs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run: timeout 120 ./pipe_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ITERS = 4096;

enum { OP_HMMA_BF16 = 0, OP_HMMA_TF32, OP_FMNMX, OP_FMNMX_NAN, OP_FMNMX3, OP_FMNMX3_NAN, OP_FFMA, OP_FFMA2, OP_LDSM4, OP_CVT_BF16X2, N_OPS };
static const char* NAMES[N_OPS] = {"mma.sync m16n8k16 bf16", "mma.sync m16n8k8 tf32", "min.f32", "min.NaN.f32", "min.f32 (3-input)",
                                   "min.NaN.f32 (3-input)", "fma.rn.f32", "fma.rn.f32x2", "ldmatrix.x4", "cvt.rn.bf16x2.f32"};

template <int OP>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cycles, float seed) {
    __shared__ __align__(16) uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float acc[8][4];
    uint32_t a[4] = {0x3f803f80u ^ threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    uint32_t b[2] = {0x3f803f80u, 0x3f803f80u};
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][j] = seed + c + j + threadIdx.x;
    unsigned long long acc2[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc2[c] = ((unsigned long long)__float_as_uint(seed + c) << 32) | __float_as_uint(seed);
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    const float fb = seed * 0.5f, fc = seed * 0.25f;
    const unsigned long long pb = ((unsigned long long)__float_as_uint(fb) << 32) | __float_as_uint(fb);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (OP == OP_HMMA_BF16) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            } else if (OP == OP_HMMA_TF32) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            } else if (OP == OP_FMNMX) {
                asm volatile("min.f32 %0, %0, %1;" : "+f"(acc[c][0]) : "f"(acc[c][1]));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(acc[c][2]) : "f"(acc[c][3]));
            } else if (OP == OP_FMNMX_NAN) {
                asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(acc[c][0]) : "f"(acc[c][1]));
                asm volatile("max.NaN.f32 %0, %0, %1;" : "+f"(acc[c][2]) : "f"(acc[c][3]));
            } else if (OP == OP_FMNMX3) {
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(acc[c][0]) : "f"(acc[c][1]), "f"(fb));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(acc[c][2]) : "f"(acc[c][3]), "f"(fc));
            } else if (OP == OP_FMNMX3_NAN) {
                asm volatile("min.NaN.f32 %0, %0, %1, %2;" : "+f"(acc[c][0]) : "f"(acc[c][1]), "f"(fb));
                asm volatile("max.NaN.f32 %0, %0, %1, %2;" : "+f"(acc[c][2]) : "f"(acc[c][3]), "f"(fc));
            } else if (OP == OP_FFMA) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[c][0]) : "f"(fb), "f"(fc));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[c][1]) : "f"(fb), "f"(fc));
            } else if (OP == OP_FFMA2) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc2[c]) : "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc2[(c + 4) & 7]) : "l"(pb));
            } else if (OP == OP_LDSM4) {
                uint32_t r0, r1, r2, r3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr + (uint32_t)((c * 512 + it * 16) & 8191)));
                a[0] ^= r0; a[1] ^= r1; a[2] ^= r2; a[3] ^= r3;
            } else if (OP == OP_CVT_BF16X2) {
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(acc[c][0]), "f"(acc[c][1]));
                a[c & 3] ^= r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(acc[c][2]), "f"(acc[c][3]));
                a[(c + 1) & 3] ^= r;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[c][j];
        s += __uint_as_float((uint32_t)(acc2[c] >> 32)) + __uint_as_float((uint32_t)acc2[c]);
    }
    s += __uint_as_float(a[0] ^ a[1] ^ a[2] ^ a[3]);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int warps, int per_iter) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, (size_t)sms * warps * 32 * sizeof(float)));
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    probe<OP><<<sms, warps * 32>>>(out, cyc, 1.0f);
    CK(cudaDeviceSynchronize());
    probe<OP><<<sms, warps * 32>>>(out, cyc, 1.0f);
    CK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(sms * sizeof(long long));
    CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double instr = (double)warps * ITERS * 8 * per_iter;
    printf("%-26s warps/SM=%2d : %8.3f warp-instr/clk/SM  (%6.2f clk per warp-instr per SM)\n", NAMES[OP], warps, instr / mx, mx / instr);
    free(h);
    CK(cudaFree(out));
    CK(cudaFree(cyc));
}

int main() {
    for (int warps : {4, 8, 16, 32}) {
        run<OP_HMMA_BF16>(warps, 1);
        run<OP_HMMA_TF32>(warps, 1);
    }
    for (int warps : {8, 32}) {
        run<OP_FMNMX>(warps, 2);
        run<OP_FMNMX_NAN>(warps, 2);
        run<OP_FMNMX3>(warps, 2);
        run<OP_FMNMX3_NAN>(warps, 2);
        run<OP_FFMA>(warps, 2);
        run<OP_FFMA2>(warps, 2);
        run<OP_LDSM4>(warps, 1);
        run<OP_CVT_BF16X2>(warps, 2);
    }
    printf("PIPE PROBE OK\n");
    return 0;
}
