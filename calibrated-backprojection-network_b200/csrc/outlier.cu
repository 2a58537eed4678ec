// OutlierRemoval.remove_outliers (reference src/net_utils.py:1761-1806) and the validity map that precedes it
// (src/kbnet.py:899-902) -- the pre-processing step immediately before the hot path.
//
//   max_value = 10 * max(sparse_depth)                                   (global, whole batch)
//   filled    = validity <= 0 ? max_value : sparse_depth
//   m         = min over the k x k window of filled (out-of-image = max_value)
//   keep      = !(m < sparse_depth - threshold)
//   validity' = validity * keep ;  depth' = sparse_depth * validity'
//
// Pure selection / comparison arithmetic: bit-exact against ATen for NaN-free inputs (the sliding minimum propagates NaN like
// ATen's max_pool2d does; the reference's global torch.max turns the WHOLE batch into NaN when one value is NaN, which this
// kernel's ordered-integer maximum does not reproduce -- NaN depths are outside the reference's own data contract).  Two launches: a grid-stride max reduction (no host
// synchronisation, unlike the reference's `full_like(fill_value=tensor)` which reads the maximum back), then one fused
// tile kernel (separable sliding minimum in shared memory).
#include "conv_common.cuh"

namespace kbnet {

namespace {

__device__ __forceinline__ unsigned f2ord(float f) {  // order-preserving map float -> unsigned
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

__global__ void max_init_kernel(unsigned* slot) { *slot = 0u; }

__global__ void max_reduce_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ slot) {
    unsigned best = 0u;  // below every finite float's image
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned o = f2ord(__ldg(x + i));
        best = o > best ? o : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(slot, best);
}

constexpr int OT = 32;       // output tile (OT x OT)
constexpr int OR_MAX = 8;    // radius limit (kernel_size <= 17)

__global__ void __launch_bounds__(256) outlier_kernel(const float* __restrict__ depth, const float* __restrict__ validity, int N, int H, int W, int r,
                                                      float threshold, const unsigned* __restrict__ max_slot, float* __restrict__ out_depth,
                                                      float* __restrict__ out_valid) {
    __shared__ float s_f[(OT + 2 * OR_MAX) * (OT + 2 * OR_MAX)];  // filled values with halo
    __shared__ float s_h[(OT + 2 * OR_MAX) * OT];                 // horizontal minima
    const int n = blockIdx.z, y0 = blockIdx.y * OT, x0 = blockIdx.x * OT;
    const int TWH = OT + 2 * r;
    const float max_value = 10.f * ord2f(*max_slot);
    const size_t plane = (size_t)H * W;
    const float* d = depth + (size_t)n * plane;
    const float* v = validity ? validity + (size_t)n * plane : nullptr;
    for (int i = threadIdx.x; i < TWH * TWH; i += blockDim.x) {
        const int ty = i / TWH, tx = i - ty * TWH;
        const int gy = y0 - r + ty, gx = x0 - r + tx;
        float f = max_value;
        if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
            const float z = __ldg(d + (size_t)gy * W + gx);
            const float val = v ? __ldg(v + (size_t)gy * W + gx) : (z > 0.f ? 1.f : z);  // src/kbnet.py:899-902
            f = val <= 0.f ? max_value : z;
        }
        s_f[i] = f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TWH * OT; i += blockDim.x) {
        const int ty = i / OT, tx = i - ty * OT;
        const float* row = s_f + ty * TWH + tx;
        float m = row[0];
        for (int k = 1; k <= 2 * r; ++k) m = min_nan(m, row[k]);
        s_h[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OT * OT; i += blockDim.x) {
        const int ty = i / OT, tx = i - ty * OT;
        const int gy = y0 + ty, gx = x0 + tx;
        if (gy >= H || gx >= W) continue;
        float m = s_h[ty * OT + tx];
        for (int k = 1; k <= 2 * r; ++k) m = min_nan(m, s_h[(ty + k) * OT + tx]);
        const size_t o = (size_t)n * plane + (size_t)gy * W + gx;
        const float z = __ldg(depth + o);
        const float val = validity ? __ldg(validity + o) : (z > 0.f ? 1.f : z);
        const float keep = (m < z - threshold) ? 0.f : 1.f;
        const float vc = val * keep;
        out_valid[o] = vc;
        out_depth[o] = z * vc;
    }
}

}  // namespace

int outlier_removal_launch(const float* depth, const float* validity, int N, int H, int W, int ksize, float threshold, float* out_depth,
                           float* out_valid, void* workspace, cudaStream_t stream) {
    KB_CHECK_ARG(ksize >= 1 && (ksize & 1) && ksize / 2 <= OR_MAX, "outlier_removal: kernel_size=%d unsupported (odd, <= %d)", ksize, 2 * OR_MAX + 1);
    KB_CHECK_ARG(N <= 65535 && cdiv(H, OT) <= 65535, "outlier_removal: more than 65535 samples or tile rows in one launch");
    unsigned* slot = (unsigned*)workspace;
    const long long n = (long long)N * H * W;
    max_init_kernel<<<1, 1, 0, stream>>>(slot);
    const int blocks = (int)((n + 256 * 8 - 1) / (256 * 8) < 1184 ? (n + 256 * 8 - 1) / (256 * 8) : 1184);
    max_reduce_kernel<<<blocks, 256, 0, stream>>>(depth, n, slot);
    dim3 grid(cdiv(W, OT), cdiv(H, OT), N);
    outlier_kernel<<<grid, 256, 0, stream>>>(depth, validity, N, H, W, ksize / 2, threshold, slot, out_depth, out_valid);
    KB_CUDA(cudaGetLastError());
    count_launch(3);
    return 0;
}

}  // namespace kbnet
