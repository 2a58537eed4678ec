// Exact-fp32 direct 3x3 convolution for layers with very few filters (F <= 4) at full resolution -- KBNet's `output0`
// (12 -> 1, reference src/networks.py:1989 via net_utils.Conv2d, output_func='linear') -- optionally fused with the depth head
// of KBNetModel.forward (src/kbnet_model.py:181-184: sigmoid, min_d / (s + min_d / max_d)).
//
// Such a layer is 108 MAC per pixel: on the tensor-core engine it is bound by the ~45-cycle floor of a tcgen05.mma whatever
// its N (profiles/r01_umma_rate_probe.log), on the CUDA cores it is simply HBM-bound (read C planes once, write F planes).
// One CTA = 64 x 16 output pixels in 4-pixel strips; the C x 18 x 68 input tile and the weights are staged in shared memory.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"
#include "umma.cuh"

namespace kbnet {

namespace {

constexpr int ST_W = 64, ST_H = 16, SC_MAX = 32, SF_MAX = 4;
constexpr int PWd = ST_W + 8, PHt = ST_H + 2;  // tile starts at x0 - 4: a TMA box must start on a 16-byte boundary (tools/tma_probe.cu)

using namespace ptx;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}

// NTHR = 256: one 4-pixel strip per thread (any F <= 4).  NTHR = 128 (F == 1, KBNet's output0): one 2-column x 4-row patch per
// thread -- 6 tile rows x 3 LDS.64 (lanes 8 bytes apart: conflict-free) feed 72 FMAs per channel, a third of the shared-memory
// wavefronts per FMA of the strip form, whose LDS.64 at a 16-byte lane pitch were 2-way conflicted and set the pace of the kernel.
template <int F, int NTHR>
__global__ void __launch_bounds__(NTHR) conv_small_kernel(const __grid_constant__ CUtensorMap tm, int use_tma, const float* __restrict__ x, long long x_stride_n, int C, int H, int W,
                                                         const float* __restrict__ w, int act, float slope, float head_min, float head_ratio,
                                                         float* __restrict__ out, long long out_stride_n) {
    extern __shared__ __align__(128) float smem[];
    float* s_x = smem;                            // [C][PHt][PWd]   (TMA destination: 128-byte aligned)
    float* s_w = smem + C * PHt * PWd;            // [C][F][12]  (9 taps + 3 pad: three 16-byte loads per (channel, filter))
    const uint32_t bar = smem_u32(s_w + C * F * 12);  // mbarrier behind the weights (no static shared memory: the TMA
                                                                 // destination must stay 128-byte aligned)
    const int n = blockIdx.z, y0 = blockIdx.y * ST_H, x0 = blockIdx.x * ST_W;
    const size_t plane = (size_t)H * W;
    for (int i = threadIdx.x; i < C * F * 12; i += NTHR) {
        const int t = i % 12, r = i / 12, f = r % F, c = r / F;
        s_w[i] = t < 9 ? __ldg(w + ((size_t)f * C + c) * 9 + t) : 0.f;
    }
    const float* xn = x + (size_t)n * x_stride_n;
    if (use_tma) {  // one bulk tensor copy stages the whole C x 18 x 72 tile (out-of-image elements read as zero = the conv padding)
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(bar, (uint32_t)(C * PHt * PWd * 4));
            tma_load_4d(smem_u32(s_x), &tm, x0 - 4, y0 - 1, 0, n, bar);
        }
        __syncthreads();
        mbar_wait(bar, 0);
    }
    // fallback tile load (tensors TMA cannot describe): batches of 8 independent loads per thread in flight (a load -> store loop would
    // leave ONE load in flight per warp and the kernel latency-bound); PWd, PHt are compile-time constants -> mul/shift index split
    const int total = use_tma ? 0 : C * PHt * PWd;
    for (int base = threadIdx.x; base < total; base += NTHR * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * NTHR;
            v[u] = 0.f;
            if (idx < total) {
                const int r = idx / PWd, tx = idx - r * PWd;
                const int c = r / PHt, ty = r - c * PHt;
                const int gy = y0 - 1 + ty, gx = x0 - 4 + tx;
                if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) v[u] = __ldg(xn + (size_t)c * plane + (size_t)gy * W + gx);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * NTHR;
            if (idx < total) s_x[idx] = v[u];
        }
    }
    __syncthreads();
    if (F == 1 && NTHR == 128) {
        const int u = threadIdx.x & 31, rg = threadIdx.x >> 5;   // output columns 2u, 2u + 1; output rows 4 rg .. 4 rg + 3
        float acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float4* wp = reinterpret_cast<const float4*>(s_w + c * 12);
            const float4 wa = wp[0], wb = wp[1], wd = wp[2];
            const float wv[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wd.x};
            // tile column of output column j and tap kx: j + 3 + kx (the tile starts at image column x0 - 4); v[k] = tile column 2u + 2 + k
            const float* xs = s_x + (c * PHt + 4 * rg) * PWd + 2 * u + 2;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const float2 t0 = *reinterpret_cast<const float2*>(xs + r * PWd), t1 = *reinterpret_cast<const float2*>(xs + r * PWd + 2),
                             t2 = *reinterpret_cast<const float2*>(xs + r * PWd + 4);
                const float v[6] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y};
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int i = r - ky;                     // tile row 4 rg + r is tap row ky of output row 4 rg + i
                    if (i >= 0 && i < 4) {
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            acc[i][e] = fmaf(v[e + 1], wv[ky * 3 + 0], fmaf(v[e + 2], wv[ky * 3 + 1], fmaf(v[e + 3], wv[ky * 3 + 2], acc[i][e])));
                    }
                }
            }
        }
        const int gx = x0 + 2 * u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gy = y0 + 4 * rg + i;
            if (gy >= H || gx >= W) continue;
            float o[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                o[e] = acc[i][e];
                if (act == 1) o[e] = leaky(o[e], slope);
                if (act == 2) o[e] = head_min / (1.f / (1.f + expf(-o[e])) + head_ratio);
            }
            float* dst = out + (size_t)n * out_stride_n + (size_t)gy * W + gx;
            if (gx + 1 < W && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
            else {
                dst[0] = o[0];
                if (gx + 1 < W) dst[1] = o[1];
            }
        }
        return;
    }
    // 256 threads: strips of 4 pixels (16 strips per row, 16 rows)
    const int px = (threadIdx.x & 15) * 4, py = threadIdx.x >> 4;
    float acc[4][F];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int f = 0; f < F; ++f) acc[i][f] = 0.f;
    for (int c = 0; c < C; ++c) {
        const float* xs = s_x + (c * PHt + py) * PWd + px + 2;  // tile column px + 2 = image column x0 + px - 2
        float wv[F][12];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const float4* wp = reinterpret_cast<const float4*>(s_w + (c * F + f) * 12);
            const float4 a = wp[0], b = wp[1], d = wp[2];
            wv[f][0] = a.x; wv[f][1] = a.y; wv[f][2] = a.z; wv[f][3] = a.w; wv[f][4] = b.x; wv[f][5] = b.y; wv[f][6] = b.z; wv[f][7] = b.w; wv[f][8] = d.x;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 t = *reinterpret_cast<const float2*>(xs + ky * PWd + 2 * q);
                v[2 * q] = t.x; v[2 * q + 1] = t.y;
            }
            // output pixel i (image column x0 + px + i) reads tile columns px + i + 1 + kx
#pragma unroll
            for (int f = 0; f < F; ++f)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    acc[i][f] = fmaf(v[i + 1], wv[f][ky * 3 + 0], fmaf(v[i + 2], wv[f][ky * 3 + 1], fmaf(v[i + 3], wv[f][ky * 3 + 2], acc[i][f])));
        }
    }
    const int gy = y0 + py, gx = x0 + px;
    if (gy >= H) return;
#pragma unroll
    for (int f = 0; f < F; ++f) {
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = acc[i][f];
            if (act == 1) o[i] = leaky(o[i], slope);
            if (act == 2) o[i] = head_min / (1.f / (1.f + expf(-o[i])) + head_ratio);  // depth head: min_d / (sigmoid(o) + min_d / max_d)
        }
        float* dst = out + (size_t)n * out_stride_n + (size_t)f * plane + (size_t)gy * W + gx;
        if (gx + 3 < W && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (gx + i < W) dst[i] = o[i];
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) ptr = nullptr;
    return (EncodeFn)ptr;
}

}  // namespace

bool conv_small_supported(int C, int F, int ksize, int stride) { return ksize == 3 && stride == 1 && F >= 1 && F <= SF_MAX && C >= 1 && C <= SC_MAX; }

int conv_small_launch(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act, float slope, float head_min,
                      float head_ratio, float* out, long long out_stride_n, cudaStream_t stream) {
    KB_CHECK_ARG(conv_small_supported(C, F, 3, 1), "conv(small): C=%d F=%d unsupported", C, F);
    KB_CHECK_ARG(N <= 65535 && cdiv(H, ST_H) <= 65535, "conv_small: more than 65535 samples or tile rows in one launch");
    dim3 grid(cdiv(W, ST_W), cdiv(H, ST_H), N);
    const size_t smem = (size_t)(C * F * 12 + C * PHt * PWd) * sizeof(float) + 16;  // sized for the real C: several CTAs per SM overlap load and math
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    int use_tma = 0;
    if (((uintptr_t)x & 15) == 0 && (W & 3) == 0 && (x_stride_n & 3) == 0) {
        static EncodeFn enc = get_encode();
        if (enc) {
            const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
            const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)x_stride_n * 4};
            const cuuint32_t box[4] = {(cuuint32_t)PWd, (cuuint32_t)PHt, (cuuint32_t)C, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            use_tma = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
    }
#define KB_SMALL(FF, NTHR)                                                                                                                    \
    do {                                                                                                                                      \
        KB_CUDA(cudaFuncSetAttribute(conv_small_kernel<FF, NTHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        conv_small_kernel<FF, NTHR><<<grid, NTHR, smem, stream>>>(tm, use_tma, x, x_stride_n, C, H, W, w, act, slope, head_min, head_ratio, out, out_stride_n);   \
    } while (0)
    static const bool strips_f1 = getenv("KBNET_SMALL_STRIPS") != nullptr;   // A/B aid: the 4-pixel-strip form for F == 1 as well
    if (F == 1 && !strips_f1) KB_SMALL(1, 128);
    else if (F == 1) KB_SMALL(1, 256);
    else if (F == 2) KB_SMALL(2, 256);
    else if (F == 3) KB_SMALL(3, 256);
    else KB_SMALL(4, 256);
#undef KB_SMALL
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace kbnet
