// Shared description of a fused convolution: out = act(conv(cat_c(sources), w)).
#pragma once
#include "common.cuh"

namespace kbnet {

enum ConvSrcMode {
    SRC_DIRECT = 0,       // NCHW tensor at the conv input resolution
    SRC_UPSAMPLE = 1,     // NCHW tensor, nearest-upsampled to the conv input resolution on the fly
    SRC_COORDS_KINV = 2,  // 3 channels generated from K^-1 (ptr = [N,9]): c(v,u) = k[3c]*u + k[3c+1]*v + k[3c+2]
    SRC_XYZ_KINV = 3,     // 3 channels coords(K^-1) * z, z = aux [N,1,ceil(H/2),ceil(W/2)] sampled at (iy/2, ix/2)
    SRC_XYZ_TENSOR = 4,   // 3 channels coordinates tensor (ptr [N,3,H,W]) * z (aux), same sampling
};

struct ConvSrcDev {
    const float* ptr;
    const float* aux;
    int C, H, W;         // channels and SOURCE spatial size
    long long stride_n;  // elements between samples of ptr
    int mode;
    float scale_y, scale_x;  // SRC_UPSAMPLE: (float)H_src / H_in, (float)W_src / W_in  (torch 'nearest')
    int zW;                  // SRC_XYZ_*: row pitch of aux;  zplane = elements per sample of aux
    long long zplane;
};

constexpr int CONV_MAX_SRC = 4;

struct ConvParams {
    ConvSrcDev src[CONV_MAX_SRC];
    int n_src;
    int N, Hin, Win, Ctot;
    int ksize, stride;
    int F;
    const float* w;  // [F, Ctot, k, k]
    float* out;      // [N, *, Hout, Wout], sample pitch out_stride_n
    long long out_stride_n;
    int Hout, Wout;
    int act;
    float slope;
};

__device__ __forceinline__ float conv_fetch(const ConvSrcDev& S, int n, int c, int iy, int ix) {
    switch (S.mode) {
        case SRC_DIRECT:
            return __ldg(S.ptr + (size_t)n * S.stride_n + ((size_t)c * S.H + iy) * S.W + ix);
        case SRC_UPSAMPLE: {
            const int sy = nearest_src(iy, S.scale_y, S.H), sx = nearest_src(ix, S.scale_x, S.W);
            return __ldg(S.ptr + (size_t)n * S.stride_n + ((size_t)c * S.H + sy) * S.W + sx);
        }
        case SRC_COORDS_KINV: {
            const float* k = S.ptr + (size_t)n * 9 + 3 * c;
            return fmaf(__ldg(k), (float)ix, fmaf(__ldg(k + 1), (float)iy, __ldg(k + 2)));
        }
        case SRC_XYZ_KINV: {
            const float* k = S.ptr + (size_t)n * 9 + 3 * c;
            const float co = fmaf(__ldg(k), (float)ix, fmaf(__ldg(k + 1), (float)iy, __ldg(k + 2)));
            return co * __ldg(S.aux + (size_t)n * S.zplane + (size_t)(iy >> 1) * S.zW + (ix >> 1));
        }
        default: {  // SRC_XYZ_TENSOR
            const float co = __ldg(S.ptr + (size_t)n * S.stride_n + ((size_t)c * S.H + iy) * S.W + ix);
            return co * __ldg(S.aux + (size_t)n * S.zplane + (size_t)(iy >> 1) * S.zW + (ix >> 1));
        }
    }
}

int conv_direct_launch(const ConvParams& p, cudaStream_t stream);
// tensor-core engine (conv_umma.cu)
bool conv_umma_supported(const ConvParams& p);
int conv_umma_launch(const ConvParams& p, const void* wpack, cudaStream_t stream);
size_t conv_wpack_bytes(const int* src_channels, int n_src, int F, int ksize);
int conv_pack_weights(const float* w, const int* src_channels, int n_src, int F, int ksize, void* wpack, cudaStream_t stream);
// second-generation engine (conv_tma.cu): persistent, TMA-fed; takes the same packed weights as conv_umma.cu
bool conv_v2_supported(const ConvParams& p);
void set_autotune(int on);
bool upconv2x_supported(const ConvSrcDev& src, int F);
int upconv2x_prepare_weights(const float* w, int F, int C, float* out4, cudaStream_t stream);
int upconv2x_launch(const ConvSrcDev& src, int N, const void* const* wpack4, int F, int act, float slope, float* out, long long out_stride_n,
                    cudaStream_t stream);
int conv_v2_launch(const ConvParams& p, const void* wpack, cudaStream_t stream);
// optional operand-format side channel of the fused KB block (operand_format.cu): `image` as bf16 hi / lo parity planes
// [N][C/8][4][H/2][W/2] x 16-byte entries, `fused` as its even pixels [N][C/8][H/2][W/2]; the same forms of conv_image /
// conv_fused (even pixels) for the next block
struct KbOperandFormat {
    const void* image_hi; const void* image_lo;
    const void* fused_hi; const void* fused_lo;
    void* out_image_hi; void* out_image_lo;
    void* out_fused_hi; void* out_fused_lo;
};
bool kb_fused_supported(const float* image, long long image_sn, const float* depth, long long depth_sn, const float* fused, long long fused_sn,
                        const float* coordinates, int C_i, int C_d, int C_f, int H, int W, int F_i, int F_d, int F_f);
int kb_fused_launch(const float* image, long long image_sn, const float* depth, long long depth_sn, const float* fused, long long fused_sn,
                    const float* coordinates, const float* kinv, int N, int C_i, int C_d, int C_f, int H, int W, const float* w_proj,
                    const void* wpack_image, int F_i, const void* wpack_depth, int F_d, const void* wpack_fused, int F_f, float slope,
                    float* out_image, long long out_image_sn, float* out_depth, long long out_depth_sn, float* out_fused, long long out_fused_sn,
                    cudaStream_t stream, const KbOperandFormat* of = nullptr);
inline int conv_launch(const ConvParams& p, const void* wpack, cudaStream_t stream) {
    // no packed weights (fp32 mode), or a layer outside both tensor-core envelopes: the exact-fp32 CUDA-core kernel
    if (!wpack || (!conv_v2_supported(p) && !conv_umma_supported(p))) return conv_direct_launch(p, stream);
    return conv_v2_supported(p) ? conv_v2_launch(p, wpack, stream) : conv_umma_launch(p, wpack, stream);
}
int kb_proj_z_launch(const float* depth, long long stride_n, const float* w, int N, int C, int H, int W, int Ho, int Wo, float slope, float* z,
                     cudaStream_t stream);
int intrinsics_inverse_launch(const float* K, int N, float sx, float sy, float* kinv0, float* kinv1, cudaStream_t stream);
int outlier_removal_launch(const float* depth, const float* validity, int N, int H, int W, int ksize, float threshold, float* out_depth,
                           float* out_valid, void* workspace, cudaStream_t stream);
bool conv_small_supported(int C, int F, int ksize, int stride);
int conv_small_launch(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act, float slope, float head_min,
                      float head_ratio, float* out, long long out_stride_n, cudaStream_t stream);
bool conv_mma_supported(int C, int F, int ksize, int stride);
int conv_mma_launch(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act, float slope, float* out,
                    long long out_stride_n, cudaStream_t stream, void* of_hi = nullptr, void* of_lo = nullptr);
int to_operand_format_launch(const float* x, long long x_stride_n, int N, int C, int H, int W, void* of_hi, void* of_lo, int even_only,
                             cudaStream_t stream);
int from_operand_format_launch(const void* of_hi, const void* of_lo, int N, int C, int H, int W, float* x, long long x_stride_n, cudaStream_t stream);
int depth_u16_decode_launch(const uint16_t* z16, long long n, float* depth, float* validity, cudaStream_t stream);
int depth_u16_encode_launch(const float* depth, long long n, uint32_t* out, cudaStream_t stream);
int depth_head_launch(const float* x, long long n, double min_d, double max_d, float* y, cudaStream_t stream);
int s2d_forward(const float* x, int N, int C_in, int H, int W, const int* min_sizes, int n_min, const int* max_sizes,
                int n_max, const float* const* w_1x1, int n_conv, int n_filter, const float* w_3x3, const void* wpack, float neg_slope,
                float* out, float* pyramid, cudaStream_t stream);
size_t s2d_wpack_bytes(int C_in, const int* min_sizes, int n_min, const int* max_sizes, int n_max, int n_conv, int n_filter, float slope);
int s2d_pack_weights(const float* const* w_1x1, int n_conv, const float* w_3x3, int C_in, const int* min_sizes, int n_min,
                     const int* max_sizes, int n_max, void* wpack, cudaStream_t stream);

}  // namespace kbnet
