/*
 * kbnet_b200.h -- C ABI of libkbnet_b200.so: B200 (sm_100a) kernels for KBNet's geometric hot path.
 *
 * The reference (alexklwong/calibrated-backprojection-network) is pure Python/PyTorch and has no FFI;
 * its "operator interface" for this path is two torch.nn.Module classes.  The entry points below are
 * what a ctypes binding for those two modules (and for the convolutions either side of them) binds;
 * each cites the reference code it replaces.  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - all tensors are contiguous NCHW float32 DEVICE pointers unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call is
 *     asynchronous with respect to the host (no host synchronisation inside);
 *   - the device is the CURRENT device of the calling thread (caller does cudaSetDevice / torch.cuda.device);
 *   - return value: 0 on success, non-zero on failure (cudaError_t value, or KBNET_E_* below);
 *     kbnet_last_error() returns a thread-local message for the last failure on this thread;
 *   - thread safety: every entry point may be called concurrently from several host threads / streams
 *     (torch.nn.DataParallel replicas call forward from one thread per device).  Process-wide state is limited to
 *     (i) two relaxed launch counters (statistics), (ii) a mutex-guarded memo of tile plans keyed by layer shape
 *     (a pure function of the shape unless autotuning is on), and (iii) the kbnet_set_autotune switch, which is
 *     process-global by design: when on, the FIRST call for a new layer shape times candidate tilings and
 *     synchronises the stream once (never inside a CUDA-graph capture: settle shapes with a warm-up call first).
 */
#ifndef KBNET_B200_H_
#define KBNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KBNET_B200_VERSION 100 /* 0.1.0 */

#define KBNET_E_BADARG      10001 /* invalid argument / unsupported configuration (message says which) */
#define KBNET_E_WORKSPACE   10002 /* workspace too small */
#define KBNET_E_ALIGN       10003 /* pointer not 16-byte aligned where required */

int         kbnet_version(void);
const char* kbnet_last_error(void);
/* Number of CUDA kernels this library has launched in this process (statistics only; monotonically increasing). */
unsigned long long kbnet_launch_count(void);
/* ... of which launches of the persistent TMA-fed tcgen05 engine (plain convolutions and fused KB blocks). */
unsigned long long kbnet_tma_engine_launch_count(void);
/* Measured tile selection for the tensor-core engine (default off; environment KBNET_AUTOTUNE=1 turns it on as well).
 * When on, the FIRST call for a new layer shape times the cost model's best tile plans on the caller's tensors (every
 * candidate writes the same result) and caches the fastest; that first call synchronises `stream`.  Later calls are
 * asynchronous as usual. */
void kbnet_set_autotune(int on);

/* ------------------------------------------------------------------------------------------------
 * SparseToDensePool.forward -- replaces reference src/networks.py:2168-2196 (module ctor :2100-2166).
 *
 *   z = x[:,0];  pyramid = [minpool_s(z) for s in min_sizes] + [maxpool_s(z) for s in max_sizes]
 *   q = pyramid; for i < n_conv: q = leaky(conv1x1(q, w_1x1[i]))
 *   out = leaky(conv3x3_pad1(cat([q, x]), w_3x3))
 *
 *   x        [N, C_in, H, W]      input (channel 0 = sparse depth)
 *   min_sizes/max_sizes           HOST int arrays, odd sizes >= 3 (sizes <= 1 must be dropped by the caller,
 *                                 as the reference ctor does at :2112-2118); size <= 33
 *   w_1x1    HOST array of n_conv DEVICE pointers: w_1x1[0] is [n_filter, n_min+n_max], others [n_filter, n_filter]
 *   w_3x3    [n_filter, n_filter + C_in, 3, 3]
 *   wpack    optional (may be NULL): the weights pre-packed by kbnet_s2d_pack_weights() into the shared-memory image the
 *            fused kernel loads with one bulk copy (saves every CTA re-deriving it from the raw tensors).  Must be NULL
 *            when kbnet_s2d_wpack_bytes() returns 0 for the configuration.  Re-pack whenever the weight values change.
 *   out      [N, n_filter, H, W]
 *   pyramid  optional (may be NULL) [N, n_min+n_max, H, W]: the pool pyramid, for bit-exact checks
 * Supported: n_filter == 8, 1 <= n_conv <= 8, 1 <= n_min+n_max <= 16, 1 <= C_in <= 8.
 * Fast path (TMA tile staging, tensor-core 3x3): C_in <= 2, distinct sizes within each list, 0 <= neg_slope <= 1;
 * other configurations run an exact-fp32 CUDA-core kernel with the same fusion.
 */
int kbnet_s2d_forward(const float* x, int N, int C_in, int H, int W,
                      const int* min_sizes, int n_min, const int* max_sizes, int n_max,
                      const float* const* w_1x1, int n_conv, int n_filter, const float* w_3x3, const void* wpack,
                      float neg_slope, float* out, float* pyramid, void* stream);

/* Size of the packed S2D weight image for this configuration, or 0 if the configuration is outside the fast path
 * (then pass wpack = NULL).  `wpack` must be 16-byte aligned device memory. */
size_t kbnet_s2d_wpack_bytes(int C_in, const int* min_sizes, int n_min, const int* max_sizes, int n_max,
                             int n_conv, int n_filter, float neg_slope);
int    kbnet_s2d_pack_weights(const float* const* w_1x1, int n_conv, const float* w_3x3, int C_in,
                              const int* min_sizes, int n_min, const int* max_sizes, int n_max, void* wpack, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CalibratedBackprojectionBlock.forward -- replaces reference src/net_utils.py:1343-1371
 * (ctor :1297-1341), with n_convolution_* == 1 as KBNet always uses (src/kbnet_model.py:107-109).
 *
 *   conv_image = leaky(conv3x3_s2_p1(image, w_image))
 *   conv_depth = leaky(conv3x3_s2_p1(cat([depth, coords]), w_depth))
 *   z          = leaky(conv1x1(depth, w_proj))            (1 output channel)
 *   conv_fused = leaky(conv1x1_s2(cat([image, coords * z, fused?]), w_fused))
 *
 *   image [N,C_i,H,W]  depth [N,C_d,H,W]  fused [N,C_f,H,W] or NULL (then C_f must be 0); each with its own
 *   `*_stride_n` = elements between samples (>= C*H*W), so channel slices of wider buffers can be read in place
 *   coords: EITHER `coordinates` [N,3,H,W] (what the reference passes; KBNetEncoder builds it at
 *           src/networks.py:317-331) OR, when `coordinates` is NULL, `kinv` [N,9] row-major K^-1 per sample,
 *           from which coords[c](v,u) = kinv[3c]*u + kinv[3c+1]*v + kinv[3c+2] is generated in-kernel.
 *   w_image [F_i,C_i,3,3]  w_depth [F_d,C_d+3,3,3]  w_proj [1,C_d,1,1]  w_fused [F_f,C_i+3+C_f,1,1]
 *   out_* [N,F,Ho,Wo] with Ho = (H-1)/2+1, Wo = (W-1)/2+1.  out_*_stride_n = elements between samples of the
 *   output (>= F*Ho*Wo; lets the caller write straight into a channel slice of a wider concat buffer).
 *   workspace: DEVICE scratch of at least kbnet_kb_workspace_bytes(...) bytes, 256-byte aligned.
 */
size_t kbnet_kb_workspace_bytes(int N, int C_i, int C_d, int C_f, int H, int W, int F_i, int F_d, int F_f);

/* Packed weights for the tensor-core engine (tcgen05, bf16 hi/lo operand split, fp32 accumulate).
 * A conv whose input is the channel concatenation of `n_src` sources with `src_channels[i]` channels each gets its
 * [F, sum(src_channels), k, k] fp32 weight re-laid once into the exact shared-memory image the kernel streams with TMA.
 * Re-pack whenever the weight values change.  `wpack` must be 16-byte aligned device memory of kbnet_conv_wpack_bytes(). */
size_t kbnet_conv_wpack_bytes(const int* src_channels, int n_src, int F, int ksize);
int    kbnet_conv_pack_weights(const float* w, const int* src_channels, int n_src, int F, int ksize, void* wpack, void* stream);

/* wpack_image / wpack_depth / wpack_fused: packed forms of w_image ({C_i}), w_depth ({C_d, 3}) and w_fused
 * ({C_i, 3, C_f} or {C_i, 3}); pass all three to run the convolutions on the tensor cores, or all NULL for the exact-fp32
 * CUDA-core path. */
int kbnet_kb_forward(const float* image, long long image_stride_n,
                     const float* depth, long long depth_stride_n,
                     const float* fused, long long fused_stride_n,
                     const float* coordinates, const float* kinv,
                     int N, int C_i, int C_d, int C_f, int H, int W,
                     const float* w_image, int F_i, const float* w_depth, int F_d,
                     const float* w_proj, const float* w_fused, int F_f,
                     const void* wpack_image, const void* wpack_depth, const void* wpack_fused, float neg_slope,
                     float* out_image, long long out_image_stride_n,
                     float* out_depth, long long out_depth_stride_n,
                     float* out_fused, long long out_fused_stride_n,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Operand format -- an OPTIONAL activation layout between this library's own layers (no reference counterpart: the
 * reference keeps NCHW fp32 between src/networks.py:364 and the KB blocks at :368-475; this is the same data, laid out the
 * way the tensor-core engine consumes it).  A tensor of C channels (C % 8 == 0, H and W even) is two DEVICE buffers, the bf16
 * hi half and the bf16 lo half of every fp32 value (x ~= hi + lo, 16 mantissa bits), each
 *     [N][C/8][4][H/2][W/2] x 16-byte entries,   entry = 8 consecutive channels of one pixel,
 *     plane index = (y & 1) * 2 + (x & 1), position (y / 2, x / 2)                       ("parity planes")
 * -- the four pixel sets a stride-2 3x3 convolution reads with unit stride.  The EVEN-PIXEL form keeps plane 0 only,
 * [N][C/8][H/2][W/2] x 16 B: it is what `fused` needs (the KB block reads it at even pixels only, src/net_utils.py:1360-1369
 * is a stride-2 1x1 convolution).  Same bytes per value as fp32; 16-byte aligned.  kbnet_operand_format_bytes() = bytes of
 * ONE half (0 for an unsupported shape).  The fused KB block loads such sources with TMA straight into its MMA operand
 * planes (no fp32 staging ring, no conversion pass); results equal the fp32-source path (same split of the same values). */
size_t kbnet_operand_format_bytes(int N, int C, int H, int W, int even_pixels_only);
int    kbnet_to_operand_format(const float* x, long long x_stride_n, int N, int C, int H, int W, void* of_hi, void* of_lo,
                               int even_pixels_only, void* stream);
int    kbnet_from_operand_format(const void* of_hi, const void* of_lo, int N, int C, int H, int W, float* x, long long x_stride_n, void* stream);

/* Optional operand-format side channel of kbnet_kb_forward_of (every member may be NULL = not used):
 *   image_hi/lo      `image` [N,C_i,H,W] in operand format; the fp32 `image` pointer may then be NULL
 *   fused_hi/lo      `fused` [N,C_f,H,W] in the EVEN-PIXEL operand format (C_f % 8 == 0); the fp32 `fused` pointer may then be NULL
 *   out_image_hi/lo  conv_image [N,F_i,Ho,Wo] written in operand format for the next block; `out_image` may then be NULL.
 *                    With `out_image` NULL and no out_image_hi/lo, conv_image is NOT COMPUTED: KBNetEncoder's last block only
 *                    reads the shape of its conv_image (src/networks.py:476-523 feed conv5_image from conv4_fused), so the
 *                    model-level forward drops that convolution; the image is then read at even pixels only.
 *   out_fused_hi/lo  the even pixels of conv_fused written in the EVEN-PIXEL operand format IN ADDITION to the fp32 `out_fused`
 *                    (a quarter of its size: what the next block's `fused_hi/lo` wants) */
typedef struct kbnet_kb_operand_format {
    const void* image_hi; const void* image_lo;
    const void* fused_hi; const void* fused_lo;
    void* out_image_hi; void* out_image_lo;
    void* out_fused_hi; void* out_fused_lo;
} kbnet_kb_operand_format;

/* 1 when a KB block of this shape runs as the fused one-launch program, i.e. may use the operand-format side channel. */
int kbnet_kb_operand_format_supported(int C_i, int C_d, int C_f, int H, int W, int F_i, int F_d, int F_f);

/* kbnet_kb_forward with the operand-format side channel.  `of` == NULL is exactly kbnet_kb_forward.  Operand-format
 * sources need the fused one-launch program (packed weights given, C_i % 16 == 0, C_d % 16 == 0, W % 4 == 0); any other
 * configuration returns KBNET_E_BADARG -- convert back with kbnet_from_operand_format and call kbnet_kb_forward. */
int kbnet_kb_forward_of(const float* image, long long image_stride_n,
                        const float* depth, long long depth_stride_n,
                        const float* fused, long long fused_stride_n,
                        const float* coordinates, const float* kinv,
                        int N, int C_i, int C_d, int C_f, int H, int W,
                        const float* w_image, int F_i, const float* w_depth, int F_d,
                        const float* w_proj, const float* w_fused, int F_f,
                        const void* wpack_image, const void* wpack_depth, const void* wpack_fused, float neg_slope,
                        float* out_image, long long out_image_stride_n,
                        float* out_depth, long long out_depth_stride_n,
                        float* out_fused, long long out_fused_stride_n,
                        void* workspace, size_t workspace_bytes, const kbnet_kb_operand_format* of, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Generic fused convolution -- replaces net_utils.Conv2d.forward (src/net_utils.py:120-141) together with the
 * torch.cat / nearest-interpolate that feed it in KBNetEncoder.forward (src/networks.py:364-365,520-531),
 * DecoderBlock.forward / UpConv2d.forward (src/net_utils.py:1453-1487, 484-499).
 *
 *   out = act(conv_kxk_stride_pad(k/2)(cat_c([src_0, src_1, ...]), w))
 *
 * Each source is an NCHW view: `ptr`, `C` channels, spatial (H, W), `stride_n` elements between samples.
 * mode 0: source resolution == conv input resolution (H_in, W_in);
 * mode 1: source is nearest-upsampled to (H_in, W_in) on the fly (torch 'nearest' index rule).
 */
typedef struct kbnet_conv_src {
    const float* ptr;
    int          C, H, W;
    long long    stride_n;
    int          mode;
} kbnet_conv_src;

int kbnet_conv2d_forward(const kbnet_conv_src* srcs, int n_src, int N, int H_in, int W_in,
                         const float* w, const void* wpack /* NULL = exact fp32 path */, int F, int ksize, int stride,
                         int act /* 0 none, 1 leaky */, float neg_slope,
                         float* out, long long out_stride_n, void* stream);

/* Nearest-2x upsample followed by a 3x3 convolution (UpConv2d, reference src/net_utils.py:484-499 with DecoderBlock's
 * interpolate to exactly twice the size, :1472-1480), computed as FOUR 2x2 convolutions of the source -- one per output
 * parity, with row/column sums of the 3x3 weights -- on the source grid: 4 taps instead of 9 and no upsampled tensor.
 *   kbnet_upconv2x_prepare_weights: w [F,C,3,3] -> w_parity [4][F,C,2,2] (parity = 2*py + px);
 *   pack each w_parity[j] with kbnet_conv_pack_weights(w_parity + j*F*C*4, {C}, 1, F, 2, ...)  (kbnet_conv_wpack_bytes(.., ksize = 2));
 *   kbnet_upconv2x_forward: x [N,C,Hs,Ws] -> out [N,F,2Hs,2Ws] = act(conv3x3(upsample2x(x), w)).
 * kbnet_upconv2x_supported() == 0 (W % 4 != 0, unaligned pointer) -> use kbnet_conv2d_forward with a mode-1 source instead. */
int kbnet_upconv2x_supported(const float* x, int C, int Hs, int Ws, long long stride_n, int F);
int kbnet_upconv2x_prepare_weights(const float* w, int F, int C, float* w_parity, void* stream);
int kbnet_upconv2x_forward(const float* x, int N, int C, int Hs, int Ws, long long stride_n, const void* const* wpack_parity, int F,
                           int act, float neg_slope, float* out, long long out_stride_n, void* stream);

/* K^-1 for the KB layers -- replaces torch.inverse + the scale_intrinsics closure in KBNetEncoder.forward
 * (src/networks.py:317-352): kinv0 = inverse(K), kinv1 = inverse(K * [[sx,1,sx],[1,sy,sy],[1,1,1]]) (may be NULL).
 * K, kinv0, kinv1: [N,9] row-major 3x3 per sample.  One launch instead of ~12 LU / triangular-solve launches per inverse. */
int kbnet_intrinsics_inverse(const float* K, int N, float scale_x, float scale_y, float* kinv0, float* kinv1, void* stream);

/* The last decoder convolution fused with the output head: y = min_d / (sigmoid(conv3x3_pad1(x, w)) + min_d / max_d)
 * (`output0`, a linear-output Conv2d, reference src/networks.py:1855-1989, followed by src/kbnet_model.py:181-184).
 * x [N,C,H,W] (C <= 32), w [F,C,3,3] (F <= 4), y [N,F,H,W]; exact fp32 on the CUDA cores -- the layer is HBM-bound. */
int kbnet_conv3x3_depth_head_forward(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F,
                                     double min_depth, double max_depth, float* y, long long y_stride_n, void* stream);

/* Conv2d with a 3x3 kernel, stride 1 and FEW channels (C <= 4 and F <= 48, or C <= 16 and F <= 16): KBNet's full-resolution
 * layers conv0_image 3 -> 48, conv0_depth 8 -> 16 (reference src/networks.py:364-365) and deconv0.conv 12 -> 12
 * (src/net_utils.py:1482-1487).  y = act(conv3x3_pad1(x, w)); act = 0 linear, 1 LeakyReLU(neg_slope), 0 <= neg_slope <= 1.
 * x [N,C,H,W] and y [N,F,H,W] fp32, either may be a channel slice of a wider NCHW buffer (*_stride_n).  Warp-level tensor
 * path (mma.sync bf16, 3-term operand split, fp32 accumulate: ~1e-5 relative) -- with so few filters the tcgen05 engine
 * behind kbnet_conv2d_forward is bound by its per-MMA issue cost.  kbnet_conv3x3_small_supported returns 1 for the shapes taken. */
int kbnet_conv3x3_small_supported(int C, int F, int ksize, int stride);
int kbnet_conv3x3_small_forward(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act,
                                float neg_slope, float* y, long long y_stride_n, void* stream);
/* ... the same with an operand-format copy of the result (F % 8 == 0) for a following KB block; `y` may be NULL when only that
 * copy is wanted (conv0_image feeds calibrated_backprojection1 only, src/networks.py:364-373). */
int kbnet_conv3x3_small_forward_of(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act,
                                   float neg_slope, float* y, long long y_stride_n, void* of_hi, void* of_lo, void* stream);

/* Element-wise half of the reference's on-disk depth format -- a 16-bit PNG holding depth * 256 (src/data_utils.py:88-167;
 * the PNG container itself is host work, kbnet_b200/io.py).
 * decode: depth = u16 / 256, values <= 0 -> 0 (load_depth, :137-141); validity (may be NULL) = 1 where depth > 0 else depth
 *         (load_depth_with_validity_map, :100-107).  z16 is a DEVICE pointer to n_elem uint16 values.
 * encode: out = uint32(depth * 256), C truncation (save_depth, :165).  Both bit-exact against numpy. */
int kbnet_depth_u16_decode(const uint16_t* z16, long long n_elem, float* depth, float* validity, void* stream);
int kbnet_depth_u16_encode(const float* depth, long long n_elem, uint32_t* out, void* stream);

/* OutlierRemoval.remove_outliers (reference src/net_utils.py:1761-1806): drop sparse points that are more than `threshold`
 * behind the nearest point of their kernel_size x kernel_size neighbourhood.  sparse_depth / validity_map / outputs are
 * [N,1,H,W]; validity_map may be NULL, then it is derived as in src/kbnet.py:899-902 (depth > 0 ? 1 : depth).  The global
 * maximum the reference reads back to the host stays on the device: `workspace` is >= 4 bytes of device scratch.
 * kernel_size odd, <= 17.  Bit-exact (selection and comparison only). */
int kbnet_outlier_removal_forward(const float* sparse_depth, const float* validity_map, int N, int H, int W, int kernel_size,
                                  float threshold, float* out_depth, float* out_validity, void* workspace, void* stream);

/* Output head of KBNetModel.forward (src/kbnet_model.py:181-184): y = min_d / (sigmoid(x) + min_d / max_d). */
int kbnet_depth_head_forward(const float* x, long long n_elem, double min_depth, double max_depth, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KBNET_B200_H_ */
