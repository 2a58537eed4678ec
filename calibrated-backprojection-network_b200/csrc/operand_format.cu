// "Operand format": an activation tensor stored the way the tcgen05 engine (conv_tma.cu) consumes it -- two tensors (the bf16 hi
// and the bf16 lo half of the fp32 value, x ~= hi + lo) of 16-byte entries, one entry = 8 consecutive channels of one pixel:
//
//     of_hi / of_lo : [N][C / 8][4][H / 2][W / 2] x 16 B        (C a multiple of 8, H and W even)
//
// with the pixels of a channel group split into the four parity planes (y & 1) * 2 + (x & 1): the stride-2 convolutions of the
// KB block read exactly such planes, so a DENSE tensor-map box over one of them lands in shared memory as a ready K-major
// operand plane (a box with traversal stride 2 over a plain [H][W] layout touches every 32-byte sector four times), and a
// consumer needs neither the raw fp32 ring nor converter warps for that source.  Plane 0 alone (the even pixels) is the
// "even-pixel" form used for `fused`, which the block reads at even pixels only: [N][C / 8][H / 2][W / 2] x 16 B.  Producers inside the library write it from their epilogues (conv_mma.cu,
// conv_tma.cu); this file holds the stand-alone conversion from NCHW fp32 (for callers that hold fp32 data, and for tests)
// and its inverse.  Same bytes per element as fp32 (2 + 2), so the format costs no HBM traffic.
#include "conv_common.cuh"
#include "umma.cuh"

namespace kbnet {

namespace {

using namespace ptx;

// i enumerates the entries in storage order: (n, cg, plane, y / 2, x / 2); `planes` = 4 (full) or 1 (even pixels only)
__device__ __forceinline__ void entry_coords(long long i, int CG, int planes, int H2, int W2, long long& n, int& cg, int& y, int& x) {
    const int xx = (int)(i % W2);
    long long r = i / W2;
    const int yy = (int)(r % H2);
    r /= H2;
    const int pl = (int)(r % planes);
    r /= planes;
    cg = (int)(r % CG);
    n = r / CG;
    y = 2 * yy + (pl >> 1);
    x = 2 * xx + (pl & 1);
}

__global__ void __launch_bounds__(256) to_operand_format_kernel(const float* __restrict__ x, long long x_stride_n, int C, int H, int W, int planes,
                                                                 long long total, uint4* __restrict__ of_hi, uint4* __restrict__ of_lo) {
    const int CG = C >> 3, HW = H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long n; int cg, py, px;
        entry_coords(i, CG, planes, H >> 1, W >> 1, n, cg, py, px);
        const float* src = x + n * x_stride_n + (long long)(cg * 8) * HW + (long long)py * W + px;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldg(src + (long long)c * HW);
        const uint2 a = split_bf16x2(v[0], v[1]), b = split_bf16x2(v[2], v[3]), c2 = split_bf16x2(v[4], v[5]), d = split_bf16x2(v[6], v[7]);
        of_hi[i] = make_uint4(a.x, b.x, c2.x, d.x);
        of_lo[i] = make_uint4(a.y, b.y, c2.y, d.y);
    }
}

__global__ void __launch_bounds__(256) from_operand_format_kernel(const uint4* __restrict__ of_hi, const uint4* __restrict__ of_lo, int C, int H, int W,
                                                                   long long total, float* __restrict__ x, long long x_stride_n) {
    const int CG = C >> 3, HW = H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long n; int cg, py, px;
        entry_coords(i, CG, 4, H >> 1, W >> 1, n, cg, py, px);
        const uint4 h = of_hi[i], l = of_lo[i];
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
        float* dst = x + n * x_stride_n + (long long)(cg * 8) * HW + (long long)py * W + px;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dst[(long long)(2 * k) * HW] = __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
            dst[(long long)(2 * k + 1) * HW] = __uint_as_float(hw[k] & 0xFFFF0000u) + __uint_as_float(lw[k] & 0xFFFF0000u);
        }
    }
}

int grid_for(long long total) {
    long long g = (total + 255) / 256;
    return (int)(g < 148LL * 16 ? (g < 1 ? 1 : g) : 148LL * 16);
}

}  // namespace

int to_operand_format_launch(const float* x, long long x_stride_n, int N, int C, int H, int W, void* of_hi, void* of_lo, int even_only,
                             cudaStream_t stream) {
    KB_CHECK_ARG(x && of_hi && of_lo && N > 0 && C > 0 && (C & 7) == 0 && H > 0 && W > 0 && !(H & 1) && !(W & 1),
                 "to_operand_format: bad argument (C must be a multiple of 8, H and W even)");
    KB_CHECK_ARG((((uintptr_t)of_hi | (uintptr_t)of_lo) & 15) == 0, "to_operand_format: outputs must be 16-byte aligned");
    KB_CHECK_ARG(x_stride_n >= (long long)C * H * W, "to_operand_format: stride_n too small");
    const int planes = even_only ? 1 : 4;
    const long long total = (long long)N * (C / 8) * planes * (H / 2) * (W / 2);
    to_operand_format_kernel<<<grid_for(total), 256, 0, stream>>>(x, x_stride_n, C, H, W, planes, total, (uint4*)of_hi, (uint4*)of_lo);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int from_operand_format_launch(const void* of_hi, const void* of_lo, int N, int C, int H, int W, float* x, long long x_stride_n, cudaStream_t stream) {
    KB_CHECK_ARG(x && of_hi && of_lo && N > 0 && C > 0 && (C & 7) == 0 && H > 0 && W > 0 && !(H & 1) && !(W & 1),
                 "from_operand_format: bad argument (C must be a multiple of 8, H and W even)");
    KB_CHECK_ARG((((uintptr_t)of_hi | (uintptr_t)of_lo) & 15) == 0, "from_operand_format: inputs must be 16-byte aligned");
    KB_CHECK_ARG(x_stride_n >= (long long)C * H * W, "from_operand_format: stride_n too small");
    const long long total = (long long)N * (C / 8) * H * W;
    from_operand_format_kernel<<<grid_for(total), 256, 0, stream>>>((const uint4*)of_hi, (const uint4*)of_lo, C, H, W, total, x, x_stride_n);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace kbnet
