// extern "C" entry points of libkbnet_b200.so (see include/kbnet_b200.h for the contract).
#include <string.h>

#include "conv_common.cuh"

namespace kbnet {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // statistics only (atomic increments)
static unsigned long long g_launches_v2 = 0;

void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
void count_launch_v2() { __atomic_fetch_add(&g_launches_v2, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static ConvSrcDev make_src(const float* ptr, const float* aux, int C, int H, int W, long long stride_n, int mode) {
    ConvSrcDev s;
    memset(&s, 0, sizeof(s));
    s.ptr = ptr; s.aux = aux; s.C = C; s.H = H; s.W = W; s.stride_n = stride_n; s.mode = mode;
    s.scale_y = s.scale_x = 1.f;
    return s;
}

static int finish_conv(ConvParams& p, int N, int Hin, int Win, const float* w, int F, int ksize, int stride, int act,
                       float slope, float* out, long long out_stride_n) {
    p.N = N; p.Hin = Hin; p.Win = Win;
    p.Ctot = 0;
    for (int i = 0; i < p.n_src; ++i) p.Ctot += p.src[i].C;
    p.ksize = ksize; p.stride = stride; p.F = F; p.w = w; p.out = out;
    const int pad = ksize / 2;
    p.Hout = (Hin + 2 * pad - ksize) / stride + 1;
    p.Wout = (Win + 2 * pad - ksize) / stride + 1;
    p.out_stride_n = out_stride_n;
    p.act = act; p.slope = slope;
    KB_CHECK_ARG(out_stride_n >= (long long)F * p.Hout * p.Wout, "conv: out_stride_n=%lld smaller than F*Ho*Wo", out_stride_n);
    return 0;
}

}  // namespace kbnet

using namespace kbnet;

extern "C" {

int kbnet_version(void) { return KBNET_B200_VERSION; }

const char* kbnet_last_error(void) { return g_err; }

unsigned long long kbnet_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
void kbnet_set_autotune(int on) { set_autotune(on); }
unsigned long long kbnet_tma_engine_launch_count(void) { return __atomic_load_n(&g_launches_v2, __ATOMIC_RELAXED); }

int kbnet_s2d_forward(const float* x, int N, int C_in, int H, int W, const int* min_sizes, int n_min, const int* max_sizes,
                      int n_max, const float* const* w_1x1, int n_conv, int n_filter, const float* w_3x3, const void* wpack,
                      float neg_slope, float* out, float* pyramid, void* stream) {
    KB_CHECK_ARG((n_min == 0 || min_sizes) && (n_max == 0 || max_sizes), "s2d: null pool size list");
    static const int none[1] = {0};
    return s2d_forward(x, N, C_in, H, W, min_sizes ? min_sizes : none, n_min, max_sizes ? max_sizes : none, n_max, w_1x1, n_conv,
                       n_filter, w_3x3, wpack, neg_slope, out, pyramid, (cudaStream_t)stream);
}

size_t kbnet_s2d_wpack_bytes(int C_in, const int* min_sizes, int n_min, const int* max_sizes, int n_max, int n_conv, int n_filter,
                             float neg_slope) {
    static const int none[1] = {0};
    return s2d_wpack_bytes(C_in, min_sizes ? min_sizes : none, n_min, max_sizes ? max_sizes : none, n_max, n_conv, n_filter, neg_slope);
}

int kbnet_s2d_pack_weights(const float* const* w_1x1, int n_conv, const float* w_3x3, int C_in, const int* min_sizes, int n_min,
                           const int* max_sizes, int n_max, void* wpack, void* stream) {
    static const int none[1] = {0};
    return s2d_pack_weights(w_1x1, n_conv, w_3x3, C_in, min_sizes ? min_sizes : none, n_min, max_sizes ? max_sizes : none, n_max, wpack,
                            (cudaStream_t)stream);
}

size_t kbnet_kb_workspace_bytes(int N, int C_i, int C_d, int C_f, int H, int W, int F_i, int F_d, int F_f) {
    (void)C_i; (void)C_d; (void)C_f; (void)F_i; (void)F_d; (void)F_f;
    const size_t Ho = (size_t)(H - 1) / 2 + 1, Wo = (size_t)(W - 1) / 2 + 1;
    return (((size_t)N * Ho * Wo * sizeof(float)) + 255) & ~(size_t)255;  // z at even pixels
}

int kbnet_kb_forward_of(const float* image, long long image_stride_n, const float* depth, long long depth_stride_n,
                        const float* fused, long long fused_stride_n, const float* coordinates, const float* kinv, int N, int C_i, int C_d, int C_f, int H, int W, const float* w_image, int F_i, const float* w_depth,
                        int F_d, const float* w_proj, const float* w_fused, int F_f, const void* wpack_image,
                        const void* wpack_depth, const void* wpack_fused, float neg_slope, float* out_image,
                        long long out_image_stride_n, float* out_depth, long long out_depth_stride_n, float* out_fused,
                        long long out_fused_stride_n, void* workspace, size_t workspace_bytes, const kbnet_kb_operand_format* of, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N == 0) return 0;
    const bool image_of = of && of->image_hi && of->image_lo, fused_of = of && of->fused_hi && of->fused_lo;
    const bool out_image_of = of && of->out_image_hi && of->out_image_lo, out_fused_of = of && of->out_fused_hi && of->out_fused_lo;
    KB_CHECK_ARG((image || image_of) && depth && w_image && w_depth && w_proj && w_fused && out_depth && out_fused, "kb: null pointer argument");
    if (of) {
        KB_CHECK_ARG((of->image_hi != nullptr) == (of->image_lo != nullptr) && (of->fused_hi != nullptr) == (of->fused_lo != nullptr) &&
                     (of->out_image_hi != nullptr) == (of->out_image_lo != nullptr) && (of->out_fused_hi != nullptr) == (of->out_fused_lo != nullptr),
                     "kb: an operand-format tensor needs both its hi and its lo half");
        const uintptr_t all = (uintptr_t)of->image_hi | (uintptr_t)of->image_lo | (uintptr_t)of->fused_hi | (uintptr_t)of->fused_lo |
                              (uintptr_t)of->out_image_hi | (uintptr_t)of->out_image_lo | (uintptr_t)of->out_fused_hi | (uintptr_t)of->out_fused_lo;
        if (all & 15) { set_error("kb: operand-format buffers must be 16-byte aligned"); return KBNET_E_ALIGN; }
        KB_CHECK_ARG(!fused_of || (C_f > 0 && C_f % 8 == 0), "kb: operand-format `fused` needs C_f %% 8 == 0 (C_f = %d)", C_f);
        KB_CHECK_ARG((!out_image_of || F_i % 8 == 0) && (!out_fused_of || F_f % 8 == 0), "kb: operand-format outputs need F %% 8 == 0");
    }
    if (fused_of && !fused) fused = depth;   // never dereferenced by the fused program; keeps the C_f / pointer contract below
    KB_CHECK_ARG((coordinates != nullptr) != (kinv != nullptr), "kb: pass exactly one of `coordinates` and `kinv`");
    KB_CHECK_ARG((fused != nullptr) == (C_f > 0), "kb: fused pointer and C_f=%d disagree", C_f);
    KB_CHECK_ARG((wpack_image != nullptr) == (wpack_depth != nullptr) && (wpack_image != nullptr) == (wpack_fused != nullptr),
                 "kb: pass all three packed weights or none");
    KB_CHECK_ARG(N >= 0 && C_i > 0 && C_d > 0 && H > 0 && W > 0 && F_i > 0 && F_d > 0 && F_f > 0, "kb: bad shape");
    if (N == 0) return 0;
    const long long plane = (long long)H * W;
    KB_CHECK_ARG((!image || image_stride_n >= C_i * plane) && depth_stride_n >= C_d * plane && (!fused || fused_of || fused_stride_n >= C_f * plane),
                 "kb: input stride_n smaller than C*H*W");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    KB_CHECK_ARG((!out_image || out_image_stride_n >= (long long)F_i * Ho * Wo) && out_depth_stride_n >= (long long)F_d * Ho * Wo &&
                 out_fused_stride_n >= (long long)F_f * Ho * Wo, "kb: output stride_n smaller than F*Ho*Wo");
    // the whole block as ONE fused launch (conv_tma.cu) whenever the tensors can be described to TMA
    if (wpack_image && kb_fused_supported(image, image_stride_n, depth, depth_stride_n, fused, fused_stride_n, coordinates, C_i, C_d, C_f, H, W,
                                          F_i, F_d, F_f))
    {
        KbOperandFormat kof;
        if (of) memcpy(&kof, of, sizeof(kof));
        static_assert(sizeof(KbOperandFormat) == sizeof(kbnet_kb_operand_format), "operand-format structs must match");
        return kb_fused_launch(image, image_stride_n, depth, depth_stride_n, fused, fused_stride_n, coordinates, kinv, N, C_i, C_d, C_f, H, W,
                               w_proj, wpack_image, F_i, wpack_depth, F_d, wpack_fused, F_f, neg_slope, out_image, out_image_stride_n,
                               out_depth, out_depth_stride_n, out_fused, out_fused_stride_n, stream, of ? &kof : nullptr);
    }
    KB_CHECK_ARG(!(image_of || fused_of || out_image_of || out_fused_of),
                 "kb: the operand-format side channel needs the fused one-launch program (packed weights, C_i %% 16 == 0, C_d %% 16 == 0, W %% 4 == 0)");
    const size_t need = kbnet_kb_workspace_bytes(N, C_i, C_d, C_f, H, W, F_i, F_d, F_f);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("kb: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return KBNET_E_WORKSPACE;
    }
    float* z = (float*)workspace;
    int rc;
    // conv_image = leaky(conv3x3 s2 (image))                                     net_utils.py:1348
    if (out_image) {
        ConvParams p; memset(&p, 0, sizeof(p));
        p.n_src = 1;
        p.src[0] = make_src(image, nullptr, C_i, H, W, image_stride_n, SRC_DIRECT);
        if ((rc = finish_conv(p, N, H, W, w_image, F_i, 3, 2, 1, neg_slope, out_image, out_image_stride_n))) return rc;
        if ((rc = conv_launch(p, wpack_image, stream))) return rc;
    }
    // conv_depth = leaky(conv3x3 s2 (cat[depth, coords]))                        net_utils.py:1351
    {
        ConvParams p; memset(&p, 0, sizeof(p));
        p.n_src = 2;
        p.src[0] = make_src(depth, nullptr, C_d, H, W, depth_stride_n, SRC_DIRECT);
        p.src[1] = coordinates ? make_src(coordinates, nullptr, 3, H, W, 3 * plane, SRC_DIRECT)
                               : make_src(kinv, nullptr, 3, H, W, 9, SRC_COORDS_KINV);
        if ((rc = finish_conv(p, N, H, W, w_depth, F_d, 3, 2, 1, neg_slope, out_depth, out_depth_stride_n))) return rc;
        if ((rc = conv_launch(p, wpack_depth, stream))) return rc;
    }
    // z = leaky(proj_depth(depth)) at even pixels                                net_utils.py:1357
    if ((rc = kb_proj_z_launch(depth, depth_stride_n, w_proj, N, C_d, H, W, Ho, Wo, neg_slope, z, stream))) return rc;
    // conv_fused = leaky(conv1x1 s2 (cat[image, coords * z, fused?]))            net_utils.py:1360-1369
    {
        ConvParams p; memset(&p, 0, sizeof(p));
        p.n_src = fused ? 3 : 2;
        p.src[0] = make_src(image, nullptr, C_i, H, W, image_stride_n, SRC_DIRECT);
        p.src[1] = coordinates ? make_src(coordinates, z, 3, H, W, 3 * plane, SRC_XYZ_TENSOR)
                               : make_src(kinv, z, 3, H, W, 9, SRC_XYZ_KINV);
        p.src[1].zW = Wo; p.src[1].zplane = (long long)Ho * Wo;
        if (fused) p.src[2] = make_src(fused, nullptr, C_f, H, W, fused_stride_n, SRC_DIRECT);
        if ((rc = finish_conv(p, N, H, W, w_fused, F_f, 1, 2, 1, neg_slope, out_fused, out_fused_stride_n))) return rc;
        if ((rc = conv_launch(p, wpack_fused, stream))) return rc;
    }
    return 0;
}

int kbnet_kb_forward(const float* image, long long image_stride_n, const float* depth, long long depth_stride_n,
                     const float* fused, long long fused_stride_n, const float* coordinates, const float* kinv, int N, int C_i, int C_d, int C_f, int H, int W, const float* w_image, int F_i, const float* w_depth,
                     int F_d, const float* w_proj, const float* w_fused, int F_f, const void* wpack_image,
                     const void* wpack_depth, const void* wpack_fused, float neg_slope, float* out_image,
                     long long out_image_stride_n, float* out_depth, long long out_depth_stride_n, float* out_fused,
                     long long out_fused_stride_n, void* workspace, size_t workspace_bytes, void* stream) {
    KB_CHECK_ARG(out_image || N == 0, "kb: null pointer argument");
    return kbnet_kb_forward_of(image, image_stride_n, depth, depth_stride_n, fused, fused_stride_n, coordinates, kinv, N, C_i, C_d, C_f, H, W,
                               w_image, F_i, w_depth, F_d, w_proj, w_fused, F_f, wpack_image, wpack_depth, wpack_fused, neg_slope, out_image,
                               out_image_stride_n, out_depth, out_depth_stride_n, out_fused, out_fused_stride_n, workspace, workspace_bytes,
                               nullptr, stream);
}

int kbnet_kb_operand_format_supported(int C_i, int C_d, int C_f, int H, int W, int F_i, int F_d, int F_f) {
    if (C_i <= 0 || C_d <= 0 || C_f < 0 || H <= 0 || W <= 0 || F_i <= 0 || F_d <= 0 || F_f <= 0) return 0;
    const float* aligned = reinterpret_cast<const float*>(uintptr_t(256));   // alignment stand-in, never dereferenced
    if ((H & 1) || (W & 1)) return 0;   // parity planes of equal size
    return kb_fused_supported(aligned, (long long)C_i * H * W, aligned, (long long)C_d * H * W, C_f ? aligned : nullptr, (long long)C_f * H * W,
                              nullptr, C_i, C_d, C_f, H, W, F_i, F_d, F_f) ? 1 : 0;
}

size_t kbnet_operand_format_bytes(int N, int C, int H, int W, int even_pixels_only) {
    if (N <= 0 || C <= 0 || (C & 7) || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return 0;
    return (size_t)N * (size_t)(C / 8) * (size_t)(even_pixels_only ? 1 : 4) * (size_t)(H / 2) * (size_t)(W / 2) * 16;
}

int kbnet_to_operand_format(const float* x, long long x_stride_n, int N, int C, int H, int W, void* of_hi, void* of_lo, int even_pixels_only,
                            void* stream) {
    if (N == 0) return 0;
    return to_operand_format_launch(x, x_stride_n, N, C, H, W, of_hi, of_lo, even_pixels_only, (cudaStream_t)stream);
}

int kbnet_from_operand_format(const void* of_hi, const void* of_lo, int N, int C, int H, int W, float* x, long long x_stride_n, void* stream) {
    if (N == 0) return 0;
    return from_operand_format_launch(of_hi, of_lo, N, C, H, W, x, x_stride_n, (cudaStream_t)stream);
}

size_t kbnet_conv_wpack_bytes(const int* src_channels, int n_src, int F, int ksize) {
    if (!src_channels || n_src < 1 || n_src > CONV_MAX_SRC || F <= 0) return 0;
    return conv_wpack_bytes(src_channels, n_src, F, ksize);
}

int kbnet_conv_pack_weights(const float* w, const int* src_channels, int n_src, int F, int ksize, void* wpack, void* stream) {
    return conv_pack_weights(w, src_channels, n_src, F, ksize, wpack, (cudaStream_t)stream);
}

int kbnet_conv2d_forward(const kbnet_conv_src* srcs, int n_src, int N, int H_in, int W_in, const float* w, const void* wpack,
                         int F, int ksize, int stride, int act, float neg_slope, float* out, long long out_stride_n, void* stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(srcs && w && out, "conv: null pointer argument");
    KB_CHECK_ARG(n_src >= 1 && n_src <= CONV_MAX_SRC, "conv: n_src=%d unsupported (1..%d)", n_src, CONV_MAX_SRC);
    KB_CHECK_ARG((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "conv: ksize=%d stride=%d unsupported", ksize, stride);
    KB_CHECK_ARG(N >= 0 && H_in > 0 && W_in > 0 && F > 0, "conv: bad shape");
    if (N == 0) return 0;
    ConvParams p; memset(&p, 0, sizeof(p));
    p.n_src = n_src;
    for (int i = 0; i < n_src; ++i) {
        const kbnet_conv_src& s = srcs[i];
        KB_CHECK_ARG(s.ptr && s.C > 0 && s.H > 0 && s.W > 0, "conv: source %d invalid", i);
        KB_CHECK_ARG(s.mode == 0 || s.mode == 1, "conv: source %d mode %d unsupported", i, s.mode);
        KB_CHECK_ARG(s.mode == 1 || (s.H == H_in && s.W == W_in), "conv: source %d is %dx%d but the conv input is %dx%d", i,
                     s.H, s.W, H_in, W_in);
        p.src[i] = make_src(s.ptr, nullptr, s.C, s.H, s.W, s.stride_n, s.mode ? SRC_UPSAMPLE : SRC_DIRECT);
        p.src[i].scale_y = (float)s.H / (float)H_in;
        p.src[i].scale_x = (float)s.W / (float)W_in;
    }
    int rc;
    if ((rc = finish_conv(p, N, H_in, W_in, w, F, ksize, stride, act, neg_slope, out, out_stride_n))) return rc;
    // a handful of filters over a few channels at full resolution: HBM-bound exact-fp32 CUDA-core kernel (conv_small.cu)
    if (n_src == 1 && srcs[0].mode == 0 && conv_small_supported(srcs[0].C, F, ksize, stride))
        return conv_small_launch(srcs[0].ptr, srcs[0].stride_n, N, srcs[0].C, H_in, W_in, w, F, act, neg_slope, 0.f, 0.f, out, out_stride_n,
                                 (cudaStream_t)stream);
    return conv_launch(p, wpack, (cudaStream_t)stream);
}

int kbnet_conv3x3_depth_head_forward(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, double min_depth,
                                     double max_depth, float* y, long long y_stride_n, void* stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(x && w && y && N > 0 && H > 0 && W > 0, "conv3x3_depth_head: bad argument");
    KB_CHECK_ARG(conv_small_supported(C, F, 3, 1), "conv3x3_depth_head: C=%d F=%d unsupported (C <= 32, F <= 4)", C, F);
    KB_CHECK_ARG(x_stride_n >= (long long)C * H * W && y_stride_n >= (long long)F * H * W, "conv3x3_depth_head: stride_n too small");
    return conv_small_launch(x, x_stride_n, N, C, H, W, w, F, 2, 0.f, (float)min_depth, (float)(min_depth / max_depth), y, y_stride_n,
                             (cudaStream_t)stream);
}

/* 3x3 stride-1 convolutions with few channels on the warp-level tensor path (csrc/conv_mma.cu) */
int kbnet_conv3x3_small_supported(int C, int F, int ksize, int stride) { return conv_mma_supported(C, F, ksize, stride) ? 1 : 0; }

int kbnet_conv3x3_small_forward(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act,
                                float neg_slope, float* y, long long y_stride_n, void* stream) {
    KB_CHECK_ARG(y || N == 0, "conv3x3_small: bad argument");
    return kbnet_conv3x3_small_forward_of(x, x_stride_n, N, C, H, W, w, F, act, neg_slope, y, y_stride_n, nullptr, nullptr, stream);
}

int kbnet_conv3x3_small_forward_of(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act,
                                   float neg_slope, float* y, long long y_stride_n, void* of_hi, void* of_lo, void* stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(x && w && (y || (of_hi && of_lo)) && N > 0 && H > 0 && W > 0, "conv3x3_small: bad argument");
    KB_CHECK_ARG(conv_mma_supported(C, F, 3, 1), "conv3x3_small: C=%d F=%d unsupported (C <= 4 and F <= 48, or C <= 16 and F <= 16)", C, F);
    KB_CHECK_ARG(act == 0 || (act == 1 && neg_slope >= 0.f && neg_slope <= 1.f), "conv3x3_small: act=%d neg_slope=%g unsupported", act, (double)neg_slope);
    KB_CHECK_ARG(x_stride_n >= (long long)C * H * W && (!y || y_stride_n >= (long long)F * H * W), "conv3x3_small: stride_n too small");
    return conv_mma_launch(x, x_stride_n, N, C, H, W, w, F, act, neg_slope, y, y_stride_n, (cudaStream_t)stream, of_hi, of_lo);
}

/* nearest-2x upsample + 3x3 conv as four source-resolution 2x2 convolutions (csrc/conv_tma.cu) */
int kbnet_upconv2x_supported(const float* x, int C, int Hs, int Ws, long long stride_n, int F) {
    ConvSrcDev s = make_src(x, nullptr, C, Hs, Ws, stride_n, SRC_DIRECT);
    return upconv2x_supported(s, F) ? 1 : 0;
}

int kbnet_upconv2x_prepare_weights(const float* w, int F, int C, float* w_parity, void* stream) {
    return upconv2x_prepare_weights(w, F, C, w_parity, (cudaStream_t)stream);
}

int kbnet_upconv2x_forward(const float* x, int N, int C, int Hs, int Ws, long long stride_n, const void* const* wpack_parity, int F, int act,
                           float neg_slope, float* out, long long out_stride_n, void* stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(x && wpack_parity && out && N > 0 && C > 0 && Hs > 0 && Ws > 0 && F > 0, "upconv: bad argument");
    KB_CHECK_ARG(out_stride_n >= 4LL * F * Hs * Ws && stride_n >= (long long)C * Hs * Ws, "upconv: stride_n too small");
    ConvSrcDev s = make_src(x, nullptr, C, Hs, Ws, stride_n, SRC_DIRECT);
    KB_CHECK_ARG(upconv2x_supported(s, F), "upconv: source cannot be described to TMA (W %% 4, alignment) -- use kbnet_conv2d_forward mode 1");
    return upconv2x_launch(s, N, wpack_parity, F, act, neg_slope, out, out_stride_n, (cudaStream_t)stream);
}

int kbnet_intrinsics_inverse(const float* K, int N, float scale_x, float scale_y, float* kinv0, float* kinv1, void* stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(K && kinv0 && N > 0, "intrinsics_inverse: bad argument");
    return intrinsics_inverse_launch(K, N, scale_x, scale_y, kinv0, kinv1, (cudaStream_t)stream);
}

int kbnet_outlier_removal_forward(const float* sparse_depth, const float* validity_map, int N, int H, int W, int kernel_size, float threshold,
                                  float* out_depth, float* out_validity, void* workspace, void* stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(sparse_depth && out_depth && out_validity && workspace && N > 0 && H > 0 && W > 0, "outlier_removal: bad argument");
    return outlier_removal_launch(sparse_depth, validity_map, N, H, W, kernel_size, threshold, out_depth, out_validity, workspace,
                                  (cudaStream_t)stream);
}

/* element-wise half of the reference's 16-bit-PNG depth I/O (csrc/depth_io.cu) */
int kbnet_depth_u16_decode(const uint16_t* z16, long long n_elem, float* depth, float* validity, void* stream) {
    KB_CHECK_ARG(n_elem >= 0 && (n_elem == 0 || (z16 && depth)), "depth_u16_decode: bad argument");
    KB_CHECK_ARG(n_elem < (1ll << 39), "depth_u16_decode: too many elements in one launch");
    return depth_u16_decode_launch(z16, n_elem, depth, validity, (cudaStream_t)stream);
}

int kbnet_depth_u16_encode(const float* depth, long long n_elem, uint32_t* out, void* stream) {
    KB_CHECK_ARG(n_elem >= 0 && (n_elem == 0 || (depth && out)), "depth_u16_encode: bad argument");
    KB_CHECK_ARG(n_elem < (1ll << 39), "depth_u16_encode: too many elements in one launch");
    return depth_u16_encode_launch(depth, n_elem, out, (cudaStream_t)stream);
}

int kbnet_depth_head_forward(const float* x, long long n_elem, double min_depth, double max_depth, float* y, void* stream) {
    KB_CHECK_ARG(x && y && n_elem >= 0, "depth_head: bad argument");
    return depth_head_launch(x, n_elem, min_depth, max_depth, y, (cudaStream_t)stream);
}

}  // extern "C"
