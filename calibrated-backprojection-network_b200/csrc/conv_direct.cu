// Exact-fp32 gather convolution (CUDA cores, FFMA).  This is the reference-grade path of the library:
// every fused input form of the hot path (channel concat, nearest upsample, K^-1 coordinate generation,
// xyz = coords * z) is expressed as a "source" that the gather reads, so the tensor-core engine
// (conv_umma.cu) can be checked against it on the GPU at full size, and shapes the tensor-core engine does
// not take still run on the device (never on the CPU).
#include "conv_common.cuh"

namespace kbnet {

template <int CO_T, int CCH>
__global__ void __launch_bounds__(128) conv_direct_kernel(const ConvParams p) {
    extern __shared__ __align__(16) float s_w[];  // [CCH][taps][CO_T]
    const int taps = p.ksize * p.ksize;
    const int pad = p.ksize / 2;
    const long long npix = (long long)p.N * p.Hout * p.Wout;
    const long long pix = (long long)blockIdx.x * 128 + threadIdx.x;
    const bool active = pix < npix;
    const int co0 = blockIdx.y * CO_T;
    int n = 0, oy = 0, ox = 0;
    if (active) {
        ox = (int)(pix % p.Wout);
        const long long t = pix / p.Wout;
        oy = (int)(t % p.Hout);
        n = (int)(t / p.Hout);
    }
    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;

    for (int cb = 0; cb < p.Ctot; cb += CCH) {
        const int cn = min(CCH, p.Ctot - cb);
        __syncthreads();
        for (int i = threadIdx.x; i < cn * taps * CO_T; i += 128) {
            const int j = i % CO_T, r = i / CO_T;
            const int t = r % taps, c = r / taps;
            const int co = co0 + j;
            s_w[i] = co < p.F ? __ldg(p.w + ((size_t)co * p.Ctot + cb + c) * taps + t) : 0.f;
        }
        __syncthreads();
        if (!active) continue;
        for (int c = 0; c < cn; ++c) {
            const int cg = cb + c;
            int si = 0, coff = 0;
            while (si + 1 < p.n_src && cg >= coff + p.src[si].C) { coff += p.src[si].C; ++si; }
            const ConvSrcDev& S = p.src[si];
            const int cl = cg - coff;
            for (int ky = 0; ky < p.ksize; ++ky) {
                const int iy = oy * p.stride - pad + ky;
                if (iy < 0 || iy >= p.Hin) continue;
                for (int kx = 0; kx < p.ksize; ++kx) {
                    const int ix = ox * p.stride - pad + kx;
                    if (ix < 0 || ix >= p.Win) continue;
                    const float v = conv_fetch(S, n, cl, iy, ix);
                    const float* wr = s_w + (c * taps + ky * p.ksize + kx) * CO_T;
#pragma unroll
                    for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
                }
            }
        }
    }
    if (!active) return;
    float* dst = p.out + (size_t)n * p.out_stride_n + (size_t)oy * p.Wout + ox;
    const size_t oplane = (size_t)p.Hout * p.Wout;
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
        const int co = co0 + j;
        if (co < p.F) dst[(size_t)co * oplane] = p.act ? leaky(acc[j], p.slope) : acc[j];
    }
}

int conv_direct_launch(const ConvParams& p, cudaStream_t stream) {
    const long long npix = (long long)p.N * p.Hout * p.Wout;
    if (npix == 0 || p.F == 0) return 0;
    const int taps = p.ksize * p.ksize;
    constexpr int CCH = 8;
    dim3 grid((unsigned)((npix + 127) / 128), 1, 1);
    if (p.F <= 8) {
        grid.y = cdiv(p.F, 8);
        conv_direct_kernel<8, CCH><<<grid, 128, CCH * taps * 8 * sizeof(float), stream>>>(p);
    } else {
        grid.y = cdiv(p.F, 16);
        conv_direct_kernel<16, CCH><<<grid, 128, CCH * taps * 16 * sizeof(float), stream>>>(p);
    }
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// z = leaky(w_proj . depth) at even pixels only (the stride-2 1x1 conv_fused never looks at the others):
// reference src/net_utils.py:1357 evaluates it at full resolution and discards 3/4 of it.
__global__ void kb_proj_z_kernel(const float* __restrict__ depth, long long stride_n, const float* __restrict__ w, int N, int C, int H, int W,
                                 int Ho, int Wo, float slope, float* __restrict__ z) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * Ho * Wo) return;
    const int ox = (int)(i % Wo);
    const long long t = i / Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    const float* d = depth + (size_t)n * stride_n + (size_t)(2 * oy) * W + 2 * ox;
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(__ldg(w + c), __ldg(d + (size_t)c * H * W), a);
    z[i] = leaky(a, slope);
}

int kb_proj_z_launch(const float* depth, long long stride_n, const float* w, int N, int C, int H, int W, int Ho, int Wo, float slope, float* z,
                     cudaStream_t stream) {
    const long long tot = (long long)N * Ho * Wo;
    if (tot == 0) return 0;
    kb_proj_z_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(depth, stride_n, w, N, C, H, W, Ho, Wo, slope, z);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

__device__ __forceinline__ void inv3x3(const double* a, float* out) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double r = 1.0 / det;
    out[0] = (float)(c00 * r); out[1] = (float)((a[2] * a[7] - a[1] * a[8]) * r); out[2] = (float)((a[1] * a[5] - a[2] * a[4]) * r);
    out[3] = (float)(c01 * r); out[4] = (float)((a[0] * a[8] - a[2] * a[6]) * r); out[5] = (float)((a[2] * a[3] - a[0] * a[5]) * r);
    out[6] = (float)(c02 * r); out[7] = (float)((a[1] * a[6] - a[0] * a[7]) * r); out[8] = (float)((a[0] * a[4] - a[1] * a[3]) * r);
}

__global__ void intrinsics_inverse_kernel(const float* __restrict__ K, int N, float sx, float sy, float* __restrict__ kinv0,
                                          float* __restrict__ kinv1) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double a[9], b[9];
    // element-wise scale in fp32 exactly as `k * scale` does in the reference (src/networks.py:352)
    const float sc[9] = {sx, 1.f, sx, 1.f, sy, sy, 1.f, 1.f, 1.f};
    for (int i = 0; i < 9; ++i) { const float k = K[n * 9 + i]; a[i] = (double)k; b[i] = (double)(k * sc[i]); }
    inv3x3(a, kinv0 + n * 9);
    if (kinv1) inv3x3(b, kinv1 + n * 9);
}

int intrinsics_inverse_launch(const float* K, int N, float sx, float sy, float* kinv0, float* kinv1, cudaStream_t stream) {
    if (N == 0) return 0;
    intrinsics_inverse_kernel<<<(N + 63) / 64, 64, 0, stream>>>(K, N, sx, sy, kinv0, kinv1);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

__global__ void depth_head_kernel(const float* __restrict__ x, long long n, float min_d, float ratio, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = 1.f / (1.f + expf(-x[i]));
    y[i] = min_d / (s + ratio);
}

int depth_head_launch(const float* x, long long n, double min_d, double max_d, float* y, cudaStream_t stream) {
    if (n == 0) return 0;
    // python-float scalars reach the tensor ops as fp32 (src/kbnet_model.py:183-184)
    depth_head_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, n, (float)min_d, (float)(min_d / max_d), y);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace kbnet
