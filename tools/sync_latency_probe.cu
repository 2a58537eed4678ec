// sync_latency_probe.cu -- what does one hand-over check cost the waiting thread on B200?
//   1. mbarrier.try_wait.parity on a phase that is ALREADY complete (the common case for a consumer that arrives late)
//   2. a plain / acquire shared-memory load of a flag word
// each as a dependent chain (the result decides a branch), single thread, idle SM and with 8 warps hammering shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_latency_probe sync_latency_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) probe(int reps, int noise, long long* out) {
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile uint32_t flag;
    __shared__ __align__(16) float junk[8192];
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        flag = 1;
    }
    for (int i = tid; i < 8192; i += 512) junk[i] = (float)i;
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");   // phase 0 completes
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            if (!ok) break;            // dependent branch, as in a wait loop
            acc += ok;
        }
        long long t1 = clock64();
        for (int i = 0; i < reps; ++i) {
            uint32_t v = flag;         // volatile LDS
            if (v == 0) break;
            acc += v;
        }
        long long t2 = clock64();
        for (int i = 0; i < reps; ++i) {
            uint32_t v;
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory");
            if (v == 0) break;
            acc += v;
        }
        long long t3 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = acc;
    } else if (noise && tid >= 64 && tid < 64 + 32 * noise) {
        // background shared-memory traffic (16-byte loads and stores, like the converter warps)
        float4 s = make_float4(0, 0, 0, 0);
        float4* j4 = reinterpret_cast<float4*>(junk);
        for (int i = 0; i < reps * 8; ++i) {
            const float4 v = j4[(tid * 7 + i * 33) & 2047];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            j4[(tid * 5 + i * 17) & 2047] = s;
        }
        if (s.x == 123.456f) out[4] = 1;
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    const int reps = 2000;
    for (int noise : {0, 4, 8}) {
        probe<<<1, 512>>>(reps, noise, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[4];
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("background warps %d: try_wait (completed phase) %.1f clk, volatile LDS %.1f clk, ld.acquire.cta %.1f clk per dependent check\n", noise,
               (double)h[0] / reps, (double)h[1] / reps, (double)h[2] / reps);
    }
    printf("SYNC LATENCY PROBE OK\n");
    return 0;
}
