// Fused SparseToDensePool forward for sm_100a -- ONE kernel for reference src/networks.py:2168-2196.
//
// Per CTA: a 16 x 64 output tile of one sample, 320 threads, ~95 KB of shared memory, 92 registers (2 CTAs per SM).
//
//   A. TMA      the sparse-depth halo tile and the second raw channel are staged into shared memory by
//               cp.async.bulk.tensor (3-D tiled tensor map over [N*C, H, W], out-of-image elements zero-filled), the
//               pre-packed weights by one cp.async.bulk, all completing on one mbarrier; the tile of the CTA that will take
//               over this slot is prefetched into L2 (cp.async.bulk.prefetch.tensor).
//   B. pass 1   vertical nested-window sweep: one thread per tile column and 6-row (R <= 8) / 9-row segment produces, for
//               every pool radius, the column min / max (FMNMX3.NAN: one instruction per radius step) -> V planes in smem.
//               The remaining warps split the raw x channels to bf16 hi/lo for the 3x3 conv in the same phase.
//   C. pass 2   horizontal sweep on 4-pixel strips, all in registers: the four windows of a strip share their core
//               (RK + 5 instead of 4 RK min / max per plane) -> 999->0 fix -> the 1x1 fusion chain (exact fp32, packed FFMA2 on
//               pixel pairs, weights broadcast from smem) -> LeakyReLU -> Q (8 channels), written to smem as bf16 hi/lo
//               "chunk planes" (16 bytes = 8 channels of one pixel; even columns then odd columns of a row).
//   D. 3x3 conv on the legacy tensor path (mma.sync m16n8k16 bf16, fp32 accumulate) with the 3-term operand split
//               a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (16 mantissa bits per operand, ~1e-5 relative).  N = 8 filters is
//               exactly one n8 tile; K = 9 taps x 8 Q channels + 3 rows x (3 dx x C_in) raw channels = 96 = 6 k16 steps.
//               A fragments come from the chunk planes with ldmatrix (a tap is a row-address shift, no im2col), are reused
//               for the three output rows a Q row contributes to and are re-loaded in place right after their last use; the
//               B fragments live in registers.  Epilogue: LeakyReLU -> 8-byte NCHW stores.
//
// Why not tcgen05 here: with N = 8 a tcgen05.mma is bounded by streaming its 128 x 16 A tile from shared memory
// (>= 47 clk per 128 pixels per k16 step measured, profiles/r01_umma_probe.md) -- 3x slower than mma.sync for this shape
// (2.13 clk per m16n8k16 per SM measured, profiles/r01_pipe_probe.log).  The dense KB convolutions do use tcgen05.
//
// Pool semantics (bit-exact vs ATen, oracle/pools_np.py): empty <=> z == 0; min over non-empty values with 999 as the
// "empty" flag and a result of exactly 999 mapped to 0; out-of-image pixels ignored (+inf / -inf); NaN propagates.
// Inside the kernel the pool planes are kept in canonical order (min pools by ascending radius, then max pools by
// ascending radius); the first 1x1 weight matrix is permuted accordingly when it is packed.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "conv_common.cuh"
#include "umma.cuh"

namespace kbnet {

int s2d_generic_forward(const float* x, int N, int C_in, int H, int W, const int* min_sizes, int n_min, const int* max_sizes,
                        int n_max, const float* const* w_1x1, int n_conv, int n_filter, const float* w_3x3, float neg_slope,
                        float* out, float* pyramid, cudaStream_t stream);

namespace {

using namespace ptx;
typedef unsigned long long u64;

constexpr int NF = 8;            // n_filter of the fast path
constexpr int TH = 16, TW = 64;  // output tile (352 = 22 x 16, 480 = 30 x 16, 1216 = 19 x 64, 640 = 10 x 64: no partial tiles)
constexpr int PH = TH + 2;       // conv-input rows   y0-1 .. y0+16  (18 x 17 = 306 strips: one round of 320 threads)
// pass-1 rows per work item: three 6-row items per tile column leave two warps for the raw-channel split at R <= 8 (252 items);
// at R = 16 (200 columns) two 9-row items per column do
__host__ __device__ constexpr int lv_of(int R) { return R <= 8 ? 6 : 9; }
constexpr int PW = TW + 4;       // conv-input cols   x0-2 .. x0+65  (17 strips of 4)
constexpr int XP = PW + 4;       // smem pitch of the raw-channel tile: TMA boxes start at x0-4 (16-byte aligned coordinate)
constexpr int XCP = PW + 4;      // pitch (32-bit words) of a packed raw-channel plane row: words PW .. PW + 3 are zero (unused mma k-slots read them)
constexpr int QROW = PW;         // Q plane row: 34 even-column entries, then 34 odd-column entries (16 bytes each)
constexpr int NT = 320;
constexpr int CONV_WARPS = 8;    // phase D: 4 column blocks x 2 row groups
constexpr int LH = 4;            // pass-2 strip width
constexpr int MAX_POOL = 16;
constexpr int MAX_CONV = 8;
constexpr int SMEM_MAX = 200 * 1024;   // beyond this the generic kernel takes the configuration (shipped ones need ~100 KB: 2 CTAs / SM)
constexpr int WB_BYTES = 6 * 2 * 32 * 8;
static_assert(PH % 6 == 0 && PH % 9 == 0 && PW % LH == 0 && TH % 2 == 0, "tiling");

struct PoolSpec {            // canonical plane order: min pools by ascending radius, then max pools by ascending radius
    int n_pool, n_min;
    int radius[MAX_POOL];    // radius of canonical plane k
    int chan[MAX_POOL];      // pyramid channel (reference order) of canonical plane k
    unsigned min_mask, max_mask;  // bit r set <=> a min / max pool of radius r exists
    int rmax_min, rmax_max;
};

struct WeightSrc {  // raw weight tensors + the canonical plane -> pyramid channel map
    const float* w1x1[MAX_CONV];
    const float* w3x3;
    int n_conv, n_pool, C_in;
    int chan[MAX_POOL];
};

struct FastParams {
    const float* x;
    float* out;
    float* pyramid;
    int N, C_in, H, W;
    PoolSpec ps;
    WeightSrc ws;
    int n_conv;
    const uint8_t* wpack;  // pre-packed smem image of the weights (s2d_pack_weights), or nullptr: stage from the raw tensors
    float slope;
    int tiles_x, tiles_y;
    int use_tma;
    int n_sm;            // SMs of the device (look-ahead distance of the L2 prefetch)
    long long* stamps;   // KBNET_S2D_STAMPS builds: [cta][6] clock64 at start, data landed, pass 1, pass 2, conv done, + SM id
};
#ifdef KBNET_S2D_STAMPS
#define S2D_STAMP(k)                                                                                                              \
    do {                                                                                                                          \
        if (p.stamps != nullptr && threadIdx.x == 0)                                                                              \
            p.stamps[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 6 + (k)] = clock64();              \
    } while (0)
#else
#define S2D_STAMP(k) do { } while (0)
#endif

__host__ __device__ inline int w1_bytes(int n_pool, int n_conv) { return (((NF * n_pool + (n_conv - 1) * NF * NF) * 4) + 15) & ~15; }

struct Layout {  // byte offsets into dynamic shared memory (same function on host and device)
    int w1, wB, xch, xcl, v, z, x1, qhi, qlo, total;
};
__host__ __device__ inline Layout make_layout(int R, int n_pool, int n_conv) {
    const int ZW = PW + 2 * R, ZH = PH + 2 * R, ZP = ZW + 4;
    Layout L;
    int o = 16;  // [0,8): mbarrier
    L.w1 = o; o += w1_bytes(n_pool, n_conv);
    L.wB = o; o += WB_BYTES;
    o = (o + 15) & ~15;
    L.xch = o; o += PH * XCP * 4;
    L.xcl = o; o += PH * XCP * 4;
    o = (o + 15) & ~15;
    L.v = o; o += n_pool * PH * ZW * 4;
    o = (o + 127) & ~127;
    L.z = o;                                       // union: {z tile, x1 tile} (dead after phase B)  /  {Q hi, Q lo}
    L.x1 = (L.z + ZH * ZP * 4 + 127) & ~127;
    const int end_in = L.x1 + PH * XP * 4;
    L.qhi = o;
    L.qlo = L.qhi + PH * PW * 16;
    const int end_q = L.qlo + PH * PW * 16;
    L.total = end_in > end_q ? end_in : end_q;
    return L;
}

// compile-time helpers for the specialised pool configurations
__host__ __device__ constexpr int cpopc(unsigned m) { return m ? (int)(m & 1u) + cpopc(m >> 1) : 0; }
__host__ __device__ constexpr int nth_bit(unsigned m, int k) {  // position of the k-th (0-based) set bit
    int pos = 0;
    for (; pos < 32; ++pos)
        if ((m >> pos) & 1u) {
            if (k == 0) return pos;
            --k;
        }
    return 0;
}
__host__ __device__ constexpr int top_bit(unsigned m) { int p = 0; for (int i = 0; i < 32; ++i) if ((m >> i) & 1u) p = i; return p; }

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// w * (x0, x1) + (c0, c1): ptxas turns the (w, w) pair into the scalar-broadcast operand form of FFMA2 (no register copy)
__device__ __forceinline__ u64 ffma2s(float w, u64 x, u64 c) { return ffma2(pack2(w, w), x, c); }
__device__ __forceinline__ u64 fmul2s(float w, u64 x) { return fmul2(pack2(w, w), x); }
// LeakyReLU for 0 <= slope <= 1:  max(v, slope * v)  (NaN stays NaN: both operands are NaN)
__device__ __forceinline__ float leaky01(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ u64 leaky01x2(u64 v, float slope) {
    float a, b, sa, sb;
    unpack2(v, a, b);
    unpack2(fmul2s(slope, v), sa, sb);
    return pack2(fmaxf(a, sa), fmaxf(b, sb));
}
// two fp32 -> packed bf16 (a in the low half), round to nearest
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// (a, b) -> hi pair by truncation (one PRMT), lo pair = rn(residual); 16 mantissa bits in total
__device__ __forceinline__ void split_trunc(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    hi = __byte_perm(ua, ub, 0x7632);
    lo = cvt_bf16x2(a - __uint_as_float(ua & 0xFFFF0000u), b - __uint_as_float(ub & 0xFFFF0000u));
}
// same with a rounded hi part (safe for arbitrary NaN payloads of raw inputs)
__device__ __forceinline__ void split_rn(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = cvt_bf16x2(a, b);
    lo = cvt_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// The innermost coordinate of a tiled TMA load must be a multiple of 16 bytes (checked on hardware, tools/tma_probe.cu:
// an unaligned coordinate raises "illegal instruction"); negative / out-of-range coordinates are fine and read as zero.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

// L2 prefetch of a tile's TMA box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- weight image: [w1 : 1x1 weights as plain floats, layer 0 as [canonical plane][j], layers >= 1 as [k][j]]
//                    [wB : 3x3 weights as per-lane mma B fragments [kstep = dy*2 + half][hi|lo][lane] -> (b0, b1)]
//   half 0: k 0-7 = Q channels at dx 0, k 8-15 = Q channels at dx 1;  half 1: k 0-7 = Q at dx 2, k 8-15 = raw x slots
//   (slot s = c + C_in * dx, s < 3 * C_in).  B fragment: n = lane / 4, b0 holds k = 2t, 2t+1, b1 holds k = 2t+8, 2t+9.
__device__ __forceinline__ void build_weight_image(const WeightSrc& ws, uint8_t* dst, int tid, int nthreads) {
    float* w1 = reinterpret_cast<float*>(dst);
    const int np = ws.n_pool;
    for (int i = tid; i < NF * np; i += nthreads) {
        const int k = i / NF, j = i - k * NF;
        w1[i] = __ldg(ws.w1x1[0] + j * np + ws.chan[k]);
    }
    for (int l = 1; l < ws.n_conv; ++l)
        for (int i = tid; i < NF * NF; i += nthreads) {
            const int k = i >> 3, j = i & 7;
            w1[NF * np + (l - 1) * NF * NF + i] = __ldg(ws.w1x1[l] + j * NF + k);
        }
    uint32_t* wB = reinterpret_cast<uint32_t*>(dst + w1_bytes(np, ws.n_conv));
    const int ctot = NF + ws.C_in;
    for (int i = tid; i < 6 * 32 * 2; i += nthreads) {
        const int reg = i & 1, lane = (i >> 1) & 31, ks = i >> 6;
        const int nn = lane >> 2, t = lane & 3;
        const int dy = ks >> 1, half = ks & 1;
        float w[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int slot = 2 * t + e;  // within the 8-slot chunk; reg selects the chunk
            int c = -1, dx = 0;
            if (reg == 0 || half == 0) {  // a Q chunk
                c = slot;
                dx = half == 0 ? reg : 2;
            } else if (slot < 3 * ws.C_in) {
                c = NF + slot % ws.C_in;
                dx = slot / ws.C_in;
            }
            w[e] = c >= 0 ? __ldg(ws.w3x3 + (((size_t)nn * ctot + c) * 3 + dy) * 3 + dx) : 0.f;
        }
        uint32_t hi, lo;
        split_rn(w[0], w[1], hi, lo);
        wB[((ks * 2 + 0) * 32 + lane) * 2 + reg] = hi;
        wB[((ks * 2 + 1) * 32 + lane) * 2 + reg] = lo;
    }
}

__global__ void s2d_pack_weights_kernel(const WeightSrc ws, uint8_t* dst) { build_weight_image(ws, dst, threadIdx.x, blockDim.x); }

// ---- pass 1: vertical nested-window sweep of one tile column, LV rows per thread ---------------------------------------------
// MINM / MAXM >= 0: the pool configuration is a compile-time constant (bit r set <=> a pool of radius r), so every mask test,
// plane index and radius loop below folds away; -1: taken from the launch parameters (any configuration).
// Each thread loads only the rows its largest radius needs (RMIN rows either side for the min pools, RMAX for the max pools).
template <int R, bool BORDER, int MINM, int MAXM>
__device__ __forceinline__ void pass1(const FastParams& p, const float* s_z, float* s_V, int j, int seg, int gy0, int gx0) {
    constexpr int ZW = PW + 2 * R, ZP = ZW + 4;
    constexpr int LV = lv_of(R);
    constexpr int NV = LV + 2 * R;
    constexpr bool SPEC = MINM >= 0;
    const float INF = __int_as_float(0x7f800000);
    const unsigned min_mask = SPEC ? (unsigned)MINM : p.ps.min_mask, max_mask = SPEC ? (unsigned)MAXM : p.ps.max_mask;
    const int n_min = SPEC ? cpopc((unsigned)(SPEC ? MINM : 0)) : p.ps.n_min;
    constexpr int RA = SPEC ? (top_bit((unsigned)(SPEC ? MINM : 0)) > top_bit((unsigned)(SPEC ? MAXM : 0)) ? top_bit((unsigned)(SPEC ? MINM : 0))
                                                                                                       : top_bit((unsigned)(SPEC ? MAXM : 0)))
                            : R;   // largest radius actually used
    const int rbase = seg * LV;
    float f[NV];
    const float* src = s_z + rbase * ZP + j + 2;
#pragma unroll
    for (int i = R - RA; i < NV - (R - RA); ++i) f[i] = src[i * ZP];
    bool colin = true;
    if (BORDER) colin = (unsigned)(gx0 + j) < (unsigned)p.W;
    float* vdst = s_V + rbase * ZW + j;
    if (min_mask) {
        constexpr int RM = SPEC ? top_bit((unsigned)(SPEC ? MINM : 0)) : R;
        float g[NV], m[LV];
#pragma unroll
        for (int i = R - RM; i < NV - (R - RM); ++i) {
            g[i] = (f[i] == 0.f) ? 999.f : f[i];
            if (BORDER) g[i] = (colin && (unsigned)(gy0 + rbase + i) < (unsigned)p.H) ? g[i] : INF;
        }
#pragma unroll
        for (int i = 0; i < LV; ++i) m[i] = g[R + i];
#pragma unroll
        for (int r = 1; r <= RM; ++r) {
            if ((min_mask >> r) != 0u) {  // some min pool has radius >= r
#pragma unroll
                for (int i = 0; i < LV; ++i) m[i] = min_nan(m[i], min_nan(g[R + i - r], g[R + i + r]));
                if ((min_mask >> r) & 1u) {
                    float* dst = vdst + __popc(min_mask & ((1u << r) - 1u)) * (PH * ZW);
#pragma unroll
                    for (int i = 0; i < LV; ++i) dst[i * ZW] = m[i];
                }
            }
        }
    }
    if (max_mask) {
        constexpr int RX = SPEC ? top_bit((unsigned)(SPEC ? MAXM : 0)) : R;
        float m[LV];
        if (BORDER) {
#pragma unroll
            for (int i = R - RX; i < NV - (R - RX); ++i) f[i] = (colin && (unsigned)(gy0 + rbase + i) < (unsigned)p.H) ? f[i] : -INF;
        }
#pragma unroll
        for (int i = 0; i < LV; ++i) m[i] = f[R + i];
#pragma unroll
        for (int r = 1; r <= RX; ++r) {
            if ((max_mask >> r) != 0u) {
#pragma unroll
                for (int i = 0; i < LV; ++i) m[i] = max_nan(m[i], max_nan(f[R + i - r], f[R + i + r]));
                if ((max_mask >> r) & 1u) {
                    float* dst = vdst + (n_min + __popc(max_mask & ((1u << r) - 1u))) * (PH * ZW);
#pragma unroll
                    for (int i = 0; i < LV; ++i) dst[i * ZW] = m[i];
                }
            }
        }
    }
}

// horizontal window of radius rk around the 4 strip pixels (v[R + i] is pixel i): min / max are order-independent, so the
// radius steps run from rk down to 1 through a fall-through switch (no predicated-off work for small radii)
template <int R, bool IS_MIN>
__device__ __forceinline__ void hwindow(const float (&v)[LH + 2 * R], int rk, float (&m)[LH]) {
#pragma unroll
    for (int i = 0; i < LH; ++i) m[i] = v[R + i];
#define KB_STEP(r)                                                                                            \
    case r:                                                                                                   \
        if (r <= R) {                                                                                         \
            _Pragma("unroll") for (int i = 0; i < LH; ++i) {                                                  \
                const float a = v[(R + i - r) >= 0 ? (R + i - r) : 0], b = v[(R + i + r) < LH + 2 * R ? (R + i + r) : 0]; \
                m[i] = IS_MIN ? min_nan(m[i], min_nan(a, b)) : max_nan(m[i], max_nan(a, b));                  \
            }                                                                                                 \
        }
    switch (rk) {
        KB_STEP(16) KB_STEP(15) KB_STEP(14) KB_STEP(13) KB_STEP(12) KB_STEP(11) KB_STEP(10) KB_STEP(9)
        KB_STEP(8) KB_STEP(7) KB_STEP(6) KB_STEP(5) KB_STEP(4) KB_STEP(3) KB_STEP(2) KB_STEP(1)
        default: break;
    }
#undef KB_STEP
}

// one pool plane with a compile-time radius: only the 16-byte groups of the V row that the window touches are loaded
template <int R, int RK, bool IS_MIN>
__device__ __forceinline__ void hwindow_ct(const float* __restrict__ src, float (&m)[LH]) {
    constexpr int LO = (R - RK) & ~3, HI = (R + LH + RK + 3) & ~3, NVV = HI - LO, C = R - LO;   // pixel i is v[C + i]
    float v[NVV];
#pragma unroll
    for (int q = 0; q < NVV / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(src + LO + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    auto op = [](float a, float b) { return IS_MIN ? min_nan(a, b) : max_nan(a, b); };
    if (RK >= 2 && LH == 4) {
        // The four windows [C + i - RK, C + i + RK] share the core [C + 3 - RK, C + RK] (2 RK - 2 elements): RK + 5 three-input
        // min / max instructions per strip instead of 4 RK.  min / max are exact, commutative, associative and NaN-propagating
        // here, so the regrouping is bit-exact.
        float core = v[C + 3 - RK];
#pragma unroll
        for (int e = C + 4 - RK; e <= C + RK; e += 2) core = e + 1 <= C + RK ? op(core, op(v[e], v[e + 1])) : op(core, v[e]);
        const float L0 = v[C - RK], L1 = v[C - RK + 1], L2 = v[C - RK + 2], R1 = v[C + RK + 1], R2 = v[C + RK + 2], R3 = v[C + RK + 3];
        const float a = op(core, op(L1, L2)), b = op(core, op(R1, R2));
        m[0] = op(a, L0); m[1] = op(a, R1); m[2] = op(b, L2); m[3] = op(b, R3);
    } else {
#pragma unroll
        for (int i = 0; i < LH; ++i) m[i] = v[C + i];
#pragma unroll
        for (int r = 1; r <= RK; ++r) {
#pragma unroll
            for (int i = 0; i < LH; ++i) m[i] = op(m[i], op(v[C + i - r], v[C + i + r]));
        }
    }
}

// ---- C. pass 2 (horizontal) -> pyramid -> 1x1 chain -> Q chunk planes, one 4-pixel strip per thread -----------------------
template <int R, int MINM, int MAXM, int NCONV, bool BORDER, bool PYR>
__device__ __forceinline__ void pass2(const FastParams& p, const float* __restrict__ s_V, const float* __restrict__ s_w1, uint4* s_qhi,
                                      uint4* s_qlo, int s, int n, int x0, int y0, int np, int n_conv) {
    constexpr int ZW = PW + 2 * R;
    constexpr bool SPEC = MINM >= 0;
    constexpr int NS = PW / LH;
    const float slope = p.slope;
    const int n_min = SPEC ? cpopc((unsigned)(SPEC ? MINM : 0)) : p.ps.n_min;
    const int pr = s / NS, c0 = (s - pr * NS) * LH;
    const int gy = y0 - 1 + pr, gxb = x0 - 2 + c0;
    const int H = p.H, W = p.W;
    const bool want_pyr = PYR && pr >= 1 && pr <= TH && gy < H;
    u64 acc[NF][2];
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[j][0] = acc[j][1] = 0ull;
    const float* vrow = s_V + pr * ZW + c0;
    // (999 -> 0) -> optional pyramid output -> first 1x1 layer
    auto consume = [&](int k, bool is_min, float (&m)[LH]) {
        if (is_min) {
#pragma unroll
            for (int i = 0; i < LH; ++i) m[i] = (m[i] == 999.f) ? 0.f : m[i];
        }
        if (PYR && want_pyr) {
            float* dst = p.pyramid + ((size_t)n * np + p.ps.chan[k]) * ((size_t)H * W) + (size_t)gy * W;
#pragma unroll
            for (int i = 0; i < LH; ++i) {
                const int c = c0 + i, gx = gxb + i;
                if (c >= 2 && c < TW + 2 && gx < W) dst[gx] = m[i];
            }
        }
        const u64 m01 = pack2(m[0], m[1]), m23 = pack2(m[2], m[3]);
        const float4* wk = reinterpret_cast<const float4*>(s_w1 + k * NF);
        const float4 wa = wk[0], wb = wk[1];
        const float w[NF] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            acc[j][0] = ffma2s(w[j], m01, acc[j][0]);
            acc[j][1] = ffma2s(w[j], m23, acc[j][1]);
        }
    };
    if (SPEC) {
        constexpr unsigned MM = SPEC ? (unsigned)MINM : 0u, XM = SPEC ? (unsigned)MAXM : 0u;
        constexpr int NMIN = cpopc(MM), NMAX = cpopc(XM);
        // explicit unrolling over the (at most 5 + 5) planes: the radius must be a template argument
#define KB_PLANE(K)                                                                                        \
        if (K < NMIN) { float m[LH]; hwindow_ct<R, nth_bit(MM, K < NMIN ? K : 0), true>(vrow + K * (PH * ZW), m); consume(K, true, m); }
        KB_PLANE(0) KB_PLANE(1) KB_PLANE(2) KB_PLANE(3) KB_PLANE(4) KB_PLANE(5) KB_PLANE(6) KB_PLANE(7)
#undef KB_PLANE
#define KB_PLANE(K)                                                                                        \
        if (K < NMAX) { float m[LH]; hwindow_ct<R, nth_bit(XM, K < NMAX ? K : 0), false>(vrow + (NMIN + K) * (PH * ZW), m); consume(NMIN + K, false, m); }
        KB_PLANE(0) KB_PLANE(1) KB_PLANE(2) KB_PLANE(3) KB_PLANE(4) KB_PLANE(5) KB_PLANE(6) KB_PLANE(7)
#undef KB_PLANE
    } else {
        for (int k = 0; k < np; ++k) {
            const float* src = vrow + k * (PH * ZW);
            float v[LH + 2 * R], m[LH];
#pragma unroll
            for (int q = 0; q < (LH + 2 * R) / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(src + 4 * q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
            if (k < n_min) hwindow<R, true>(v, p.ps.radius[k], m);
            else           hwindow<R, false>(v, p.ps.radius[k], m);
            consume(k, k < n_min, m);
        }
    }
#pragma unroll
    for (int j = 0; j < NF; ++j) { acc[j][0] = leaky01x2(acc[j][0], slope); acc[j][1] = leaky01x2(acc[j][1], slope); }

#pragma unroll 1
    for (int l = 1; l < n_conv; ++l) {
        const float4* wl = reinterpret_cast<const float4*>(s_w1 + NF * np + (l - 1) * NF * NF);
        u64 nxt[NF][2];
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            const float4 wa = wl[2 * k], wb = wl[2 * k + 1];
            const float w[NF] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                if (k == 0) {
                    nxt[j][0] = fmul2s(w[j], acc[0][0]);
                    nxt[j][1] = fmul2s(w[j], acc[0][1]);
                } else {
                    nxt[j][0] = ffma2s(w[j], acc[k][0], nxt[j][0]);
                    nxt[j][1] = ffma2s(w[j], acc[k][1], nxt[j][1]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NF; ++j) { acc[j][0] = leaky01x2(nxt[j][0], slope); acc[j][1] = leaky01x2(nxt[j][1], slope); }
    }
    // Q -> bf16 hi/lo chunk planes; pixels outside the image are the zero padding of the 3x3 conv
    float q[NF][LH];
#pragma unroll
    for (int j = 0; j < NF; ++j) { unpack2(acc[j][0], q[j][0], q[j][1]); unpack2(acc[j][1], q[j][2], q[j][3]); }
    const bool row_in = (unsigned)gy < (unsigned)H;
#pragma unroll
    for (int i = 0; i < LH; ++i) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) split_trunc(q[2 * jp][i], q[2 * jp + 1][i], hi[jp], lo[jp]);
        if (BORDER) {
            const bool inb = row_in && (unsigned)(gxb + i) < (unsigned)W;
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) { hi[jp] = inb ? hi[jp] : 0u; lo[jp] = inb ? lo[jp] : 0u; }
        }
        // column c -> plane (c & 1), entry (c >> 1): the conv's mma rows are the even pixels, then the odd pixels of a block
        // (adjacent pixels per lane -> 8-byte output stores), and a strip's four stores hit 2 banks groups instead of 1
        const int e = pr * QROW + (i & 1) * (QROW / 2) + ((c0 + i) >> 1);
        s_qhi[e] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        s_qlo[e] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

template <int R, int MINM, int MAXM, int NCONV, int CIN, bool PYR>
__global__ void __launch_bounds__(NT, 2)
s2d_fused_kernel(const __grid_constant__ FastParams p, const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_x1) {
    constexpr int ZW = PW + 2 * R, ZH = PH + 2 * R, ZP = ZW + 4;
    constexpr bool SPEC = MINM >= 0;  // pool sizes, n_convolution and input_channels are compile-time constants
    extern __shared__ __align__(128) uint8_t smem[];
    const int np = SPEC ? cpopc((unsigned)(SPEC ? MINM : 0)) + cpopc((unsigned)(SPEC ? MAXM : 0)) : p.ps.n_pool;
    const int n_conv = SPEC ? NCONV : p.n_conv;
    const Layout L = make_layout(R, np, n_conv);
    float* s_w1 = reinterpret_cast<float*>(smem + L.w1);
    uint2* s_wB = reinterpret_cast<uint2*>(smem + L.wB);
    uint32_t* s_xch = reinterpret_cast<uint32_t*>(smem + L.xch);   // [row][XCP] packed bf16 hi words of the raw channels
    uint32_t* s_xcl = reinterpret_cast<uint32_t*>(smem + L.xcl);   // ... and the lo words
    float* s_V = reinterpret_cast<float*>(smem + L.v);
    float* s_z = reinterpret_cast<float*>(smem + L.z);
    float* s_x1 = reinterpret_cast<float*>(smem + L.x1);
    uint4* s_qhi = reinterpret_cast<uint4*>(smem + L.qhi);
    uint4* s_qlo = reinterpret_cast<uint4*>(smem + L.qlo);
    const uint32_t bar = smem_u32(smem);

    const int tid = threadIdx.x;
    const int n = blockIdx.z;                       // 3-D grid: no index arithmetic (two integer divisions were ~7 % of the kernel)
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;
    const int gy0 = y0 - 1 - R, gx0 = x0 - 2 - R;  // image coordinates of logical z-tile element (0, 0) (smem column 2)
    const int C_in = SPEC ? CIN : p.C_in;
    const float slope = p.slope;

    S2D_STAMP(0);
#ifdef KBNET_S2D_STAMPS
    if (p.stamps != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.stamps[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 6 + 5] = smid;
    }
#endif
    // ---- A. stage the input tile and the weights ---------------------------------------------------------------------
    const bool bulk_w = p.wpack != nullptr;
    const uint32_t wbytes = (uint32_t)(w1_bytes(np, n_conv) + WB_BYTES);
    if (tid == 0 && (p.use_tma || bulk_w)) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async_smem();
        uint32_t bytes = bulk_w ? wbytes : 0u;
        if (p.use_tma) bytes += ZH * ZP * 4 + (C_in > 1 ? PH * XP * 4 : 0);
        mbar_arrive_expect_tx(bar, bytes);
        if (p.use_tma) {
            tma_load_3d(smem_u32(s_z), &tm_z, gx0 - 2, gy0, n * C_in, bar);
            if (C_in > 1) tma_load_3d(smem_u32(s_x1), &tm_x1, x0 - 4, y0 - 1, n * C_in + 1, bar);
        }
        if (bulk_w) bulk_g2s(smem_u32(s_w1), p.wpack, wbytes, bar);
#ifndef KBNET_S2D_NO_PREFETCH
        // The CTA that will take over this slot (two CTAs per SM, hardware scheduling in launch order: ~2 x #SMs CTAs ahead)
        // starts by waiting ~2.2 k clk for its input tile (12 % of a CTA's life, profiles/r02b_s2d_cta_phases.txt): ask for that
        // tile's boxes now, so that its TMA loads hit in L2.
        if (p.use_tma) {
            const unsigned ahead = 2u * (unsigned)p.n_sm;
            unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + ahead;
            const unsigned per_n = gridDim.x * gridDim.y;
            const unsigned nn = lin / per_n;
            if (nn < gridDim.z) {
                lin -= nn * per_n;
                const unsigned ty = lin / gridDim.x, tx = lin - ty * gridDim.x;
                const int px0 = (int)tx * TW, py0 = (int)ty * TH;
                tma_prefetch_3d(&tm_z, px0 - 2 - R - 2, py0 - 1 - R, (int)nn * C_in);
                if (C_in > 1) tma_prefetch_3d(&tm_x1, px0 - 4, py0 - 1, (int)nn * C_in + 1);
            }
        }
#endif
    }
    if (!p.use_tma) {  // shapes the tensor map cannot describe (W % 4 != 0, unaligned base): plain loads, same smem image
        const float* xin = p.x + (size_t)n * C_in * plane;
        for (int idx = tid; idx < ZH * ZP; idx += NT) {
            const int i = idx / ZP, j = idx - i * ZP;
            const int gy = gy0 + i, gx = gx0 - 2 + j;
            s_z[idx] = ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) ? __ldg(xin + (size_t)gy * W + gx) : 0.f;
        }
        if (C_in > 1)
            for (int idx = tid; idx < PH * XP; idx += NT) {
                const int i = idx / XP, j = idx - i * XP;
                const int gy = y0 - 1 + i, gx = x0 - 4 + j;
                s_x1[idx] = ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) ? __ldg(xin + plane + (size_t)gy * W + gx) : 0.f;
            }
    }
    if (!bulk_w) build_weight_image(p.ws, smem + L.w1, tid, NT);
    __syncthreads();
    if (p.use_tma || bulk_w) mbar_wait(bar, 0);
    S2D_STAMP(1);

    // ---- B. pass 1 (vertical): the first warps own one (column, 7-row segment) each; the remaining warps split the raw
    //         channels of the 3x3 conv input to packed bf16 -------------------------------------------------------------
    const bool border = gy0 < 0 || gx0 < 0 || gy0 + ZH > H || gx0 + ZW > W;
    {
        constexpr int N_P1 = (PH / lv_of(R)) * ZW;                       // work items of pass 1
        constexpr int T_P1 = ((N_P1 + 31) / 32) * 32;      // ... rounded up to whole warps
        static_assert(T_P1 < NT, "pass 1 leaves no warp for the raw-channel split");
        if (tid < T_P1) {
            if (tid < N_P1) {
                const int seg = tid / ZW, j = tid - seg * ZW;
                if (border) pass1<R, true, MINM, MAXM>(p, s_z, s_V, j, seg, gy0, gx0);
                else        pass1<R, false, MINM, MAXM>(p, s_z, s_V, j, seg, gy0, gx0);
            }
        } else {
            // raw channels of the 3x3 conv input as packed bf16 words (two k-slots each):
            //   C_in == 2: (x0, x1) of this pixel;   C_in == 1: x0 of this pixel and of the next one to the right
            // words PW, PW + 1 of every row are zero: k-slots that carry no raw channel read them
            // four columns per work item: 8-byte loads of the two raw rows, 16-byte stores of the hi and the lo words
            constexpr int STEP = NT - T_P1, NG = XCP / 4;      // column groups per row
            static_assert(XCP % 4 == 0 && (R % 2) == 0, "8-byte alignment of the raw-row loads, 16-byte alignment of the word stores");
            int it = tid - T_P1, pr = it / NG, cg = it - pr * NG;
            for (; it < PH * NG; it += STEP, cg += STEP % NG, pr += STEP / NG) {
                if (cg >= NG) { cg -= NG; ++pr; }
                const int c = 4 * cg;
                float a[5], bb[4];
                const float* zr = s_z + (pr + R) * ZP + c + R + 2;     // even float offset: 8-byte aligned
                if (c + 4 <= PW) {
                    const float2 z0 = *reinterpret_cast<const float2*>(zr), z1 = *reinterpret_cast<const float2*>(zr + 2);
                    a[0] = z0.x; a[1] = z0.y; a[2] = z1.x; a[3] = z1.y;
                    a[4] = C_in > 1 ? 0.f : zr[4];
                    if (C_in > 1) {
                        const float* xr = s_x1 + pr * XP + c + 2;
                        const float2 x0v = *reinterpret_cast<const float2*>(xr), x1v = *reinterpret_cast<const float2*>(xr + 2);
                        bb[0] = x0v.x; bb[1] = x0v.y; bb[2] = x1v.x; bb[3] = x1v.y;
                    } else {
                        bb[0] = a[1]; bb[1] = a[2]; bb[2] = a[3]; bb[3] = a[4];
                    }
                } else {                                               // the zero words at the end of the row
#pragma unroll
                    for (int e = 0; e < 4; ++e) a[e] = bb[e] = 0.f;
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_rn(a[e], bb[e], hi[e], lo[e]);
                *reinterpret_cast<uint4*>(s_xch + pr * XCP + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(s_xcl + pr * XCP + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    __syncthreads();
    S2D_STAMP(2);

    // ---- C. pass 2 (horizontal) -> pyramid -> 1x1 chain -> Q chunk planes ----------------------------------------------
    {
        constexpr int NS = PW / LH;
        static_assert(PH * NS <= NT, "one strip per thread");
        if (tid < PH * NS) {
            if (border) pass2<R, MINM, MAXM, NCONV, true, PYR>(p, s_V, s_w1, s_qhi, s_qlo, tid, n, x0, y0, np, n_conv);
            else        pass2<R, MINM, MAXM, NCONV, false, PYR>(p, s_V, s_w1, s_qhi, s_qlo, tid, n, x0, y0, np, n_conv);
        }
    }
    __syncthreads();
    S2D_STAMP(3);

    // ---- D. 3x3 conv on mma.sync (bf16 x3 split), LeakyReLU, NCHW stores ---------------------------------------------
    if (tid < CONV_WARPS * 32) {
        const int warp = tid >> 5, lane = tid & 31;
        constexpr int LR = TH / 2;                 // output rows per warp
        const int cb = warp & 3, rg = warp >> 2;  // 16-column block, row group
        const int g = lane >> 2, t = lane & 3;
        uint2 bf[6][2];
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
            bf[ks][0] = s_wB[(ks * 2 + 0) * 32 + lane];
            bf[ks][1] = s_wB[(ks * 2 + 1) * 32 + lane];
        }
        // mma row r <-> block pixel m = 2r (r < 8) or 2(r - 8) + 1: the fragment registers of one lane hold ADJACENT pixels, so the
        // raw-channel words are adjacent and the output stores are 8-byte accesses.  Q column of (pixel m, tap dx) =
        // 16 cb + m + 1 + dx; columns live in an even plane and an odd plane, so the 8 rows of every ldmatrix matrix are
        // consecutive 16-byte entries.  ldmatrix row addresses: matrix mi = lane / 8 -> row half (mi & 1), k chunk (mi >> 1).
        const int mi = lane >> 3, mr = lane & 7;
        const int c_s0 = 16 * cb + 2 * mr + 1 + (mi & 1) + (mi >> 1);   // k-step half 0: chunk 0 = dx 0, chunk 1 = dx 1
        const int c_s1 = 16 * cb + 2 * mr + 1 + (mi & 1) + 2;           // k-step half 1, chunk 0: dx = 2 (x2: lanes 0-15)
        const uint32_t e_s0 = (uint32_t)((c_s0 & 1) * (QROW / 2) + (c_s0 >> 1) + LR * rg * QROW) * 16u;
        const uint32_t e_s1 = (uint32_t)((c_s1 & 1) * (QROW / 2) + (c_s1 >> 1) + LR * rg * QROW) * 16u;
        // raw-channel fragment: lane t holds k-slots (2t, 2t+1) of pixels 2g and 2g + 1 (adjacent words); a lane whose slots
        // carry nothing reads the zero words at the end of the row
        int col_x;
        uint32_t xmask = 0xFFFFFFFFu;
        if (C_in > 1) { col_x = t < 3 ? 16 * cb + 2 * g + 1 + t : PW; }
        else          { col_x = t < 2 ? 16 * cb + 2 * g + 1 + 2 * t : PW; xmask = t == 1 ? 0x0000FFFFu : 0xFFFFFFFFu; }
        const uint32_t qh = smem_u32(s_qhi), ql = smem_u32(s_qlo);
        const uint32_t* xrow_h = s_xch + LR * rg * XCP + col_x;
        const uint32_t* xrow_l = s_xcl + LR * rg * XCP + col_x;
        // epilogue addressing: c0,c1 = (pixel 2g, filters 2t, 2t+1), c2,c3 = (pixel 2g + 1, same filters)
        const int gxa = x0 + 16 * cb + 2 * g;
        float* d0 = p.out + ((size_t)n * NF + 2 * t) * plane + (size_t)(y0 + LR * rg) * W + gxa;
        const bool full = y0 + TH <= H && x0 + TW <= W && (W & 1) == 0 && ((uintptr_t)p.out & 7) == 0;   // every tile of the shipped shapes
        const bool in_a = gxa < W, in_b = gxa + 1 < W;
        float* r0 = d0;                                    // output row pointers (filters 2t and 2t + 1), advanced row by row
        float* r1 = d0 + plane;
        // Only three output rows are open at any time (row li takes tap rows li .. li + 2); the issue order walks them
        // round-robin so that consecutive HMMAs never touch the same accumulator (a dependent HMMA waits for its predecessor).
        float acc[3][4];
        // The A fragments of tap row qr + 1 are requested while the HMMAs of row qr still issue: each fragment register set is
        // reloaded right after its last use (a0l after term 1, a0h after term 2, a1l after term 4, a1h after term 5), so the
        // ldmatrix / LDS latency hides behind the remaining HMMAs without a second register set (the kernel sits at the register
        // limit of two CTAs per SM).  Before: the HMMAs' top stall was the shared-memory scoreboard (profiles/r02b_s2d_phases.txt).
        uint32_t a0h[4], a0l[4], a1h[4], a1l[4];
        auto load_a0h = [&](int qr) { ldsm_x4(qh + e_s0 + (uint32_t)(qr * QROW * 16), a0h); };
        auto load_a0l = [&](int qr) { ldsm_x4(ql + e_s0 + (uint32_t)(qr * QROW * 16), a0l); };
        auto load_a1h = [&](int qr) {
            ldsm_x2(qh + e_s1 + (uint32_t)(qr * QROW * 16), a1h[0], a1h[1]);
            a1h[2] = xrow_h[qr * XCP]; a1h[3] = xrow_h[qr * XCP + 1];
            if (C_in == 1) { a1h[2] &= xmask; a1h[3] &= xmask; }
        };
        auto load_a1l = [&](int qr) {
            ldsm_x2(ql + e_s1 + (uint32_t)(qr * QROW * 16), a1l[0], a1l[1]);
            a1l[2] = xrow_l[qr * XCP]; a1l[3] = xrow_l[qr * XCP + 1];
            if (C_in == 1) { a1l[2] &= xmask; a1l[3] &= xmask; }
        };
        load_a0h(0); load_a0l(0); load_a1h(0); load_a1l(0);
#pragma unroll
        for (int qr = 0; qr < LR + 2; ++qr) {
            if (qr < LR) {                                 // output row qr opens with tap row 0
                float (&m)[4] = acc[qr % 3];
                m[0] = m[1] = m[2] = m[3] = 0.f;
            }
#pragma unroll
            for (int term = 0; term < 6; ++term) {         // (A fragment, B fragment) pairs of one tap row
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int li = qr - dy;
                    if (li >= 0 && li < LR) {
                        const int ks = 2 * dy + (term >= 3 ? 1 : 0);
                        const uint2 bh = bf[ks][0], bl = bf[ks][1];
                        const int tt = term % 3;           // 0: a_hi b_hi   1: a_lo b_hi   2: a_hi b_lo
                        if (term < 3) {
                            if (tt == 0)      hmma(acc[li % 3], a0h, bh.x, bh.y);
                            else if (tt == 1) hmma(acc[li % 3], a0l, bh.x, bh.y);
                            else              hmma(acc[li % 3], a0h, bl.x, bl.y);
                        } else {
                            if (tt == 0)      hmma(acc[li % 3], a1h, bh.x, bh.y);
                            else if (tt == 1) hmma(acc[li % 3], a1l, bh.x, bh.y);
                            else              hmma(acc[li % 3], a1h, bl.x, bl.y);
                        }
                    }
                }
                if (qr + 1 < LR + 2) {                     // this fragment set is dead now: fetch it for the next tap row
                    if (term == 1) load_a0l(qr + 1);
                    else if (term == 2) load_a0h(qr + 1);
                    else if (term == 4) load_a1l(qr + 1);
                    else if (term == 5) load_a1h(qr + 1);
                }
            }
            if (qr >= 2) {                                 // output row qr - 2 is complete: LeakyReLU, 8-byte stores
                const int li = qr - 2;
                const float (&a)[4] = acc[li % 3];
                if (full) {
                    float v0, v1, v2, v3;
                    unpack2(leaky01x2(pack2(a[0], a[1]), slope), v0, v1);
                    unpack2(leaky01x2(pack2(a[2], a[3]), slope), v2, v3);
                    *reinterpret_cast<float2*>(r0) = make_float2(v0, v2);
                    *reinterpret_cast<float2*>(r1) = make_float2(v1, v3);
                } else if (y0 + LR * rg + li < H) {
                    if (in_a) { r0[0] = leaky01(a[0], slope); r1[0] = leaky01(a[1], slope); }
                    if (in_b) { r0[1] = leaky01(a[2], slope); r1[1] = leaky01(a[3], slope); }
                }
                r0 += W; r1 += W;
            }
        }
    }
#ifdef KBNET_S2D_STAMPS
    __syncthreads();
    S2D_STAMP(4);
#endif
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    return fn;
}

// 3-D tiled tensor map over x viewed as [N*C_in][H][W] fp32, box = {bw, bh, 1}, out-of-bounds elements read as zero
bool make_map(CUtensorMap* tm, const float* x, int NC, int H, int W, int bw, int bh) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    if ((W & 3) || ((uintptr_t)x & 15) || bw > 256 || bh > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NC};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int R, int MINM, int MAXM, int NCONV, int CIN, bool PYR>
int launch_fast_t(FastParams& p, cudaStream_t stream) {
    const Layout L = make_layout(R, p.ps.n_pool, p.n_conv);
    alignas(64) CUtensorMap tm_z, tm_x1;
    memset(&tm_z, 0, sizeof(tm_z));
    memset(&tm_x1, 0, sizeof(tm_x1));
    static const bool no_tma = getenv("KBNET_S2D_NO_TMA") != nullptr;  // debugging aid: stage the tile with plain loads
    p.use_tma = !no_tma && make_map(&tm_z, p.x, p.N * p.C_in, p.H, p.W, PW + 2 * R + 4, PH + 2 * R) &&
                (p.C_in == 1 || make_map(&tm_x1, p.x, p.N * p.C_in, p.H, p.W, XP, PH));
    auto kern = s2d_fused_kernel<R, MINM, MAXM, NCONV, CIN, PYR>;
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    p.tiles_x = cdiv(p.W, TW);
    p.tiles_y = cdiv(p.H, TH);
    {
        int dev = 0, v = 0;
        p.n_sm = 148;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) p.n_sm = v;
    }
    KB_CHECK_ARG(p.tiles_y <= 65535 && p.N <= 65535, "s2d: more than 65535 tile rows or samples in one launch");
#ifdef KBNET_S2D_STAMPS
    static const char* stamp_file = getenv("KBNET_S2D_STAMPS");   // debugging aid: per-CTA phase times (synchronises!)
    const size_t n_cta = (size_t)p.tiles_x * p.tiles_y * p.N;
    if (stamp_file) {
        KB_CUDA(cudaMalloc(&p.stamps, n_cta * 6 * sizeof(long long)));
        KB_CUDA(cudaMemsetAsync(p.stamps, 0, n_cta * 6 * sizeof(long long), stream));
    }
#endif
    kern<<<dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)p.N), NT, L.total, stream>>>(p, tm_z, tm_x1);
    KB_CUDA(cudaGetLastError());
    count_launch();
#ifdef KBNET_S2D_STAMPS
    if (stamp_file && p.stamps) {
        std::vector<long long> h(n_cta * 6);
        KB_CUDA(cudaStreamSynchronize(stream));
        KB_CUDA(cudaMemcpy(h.data(), p.stamps, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.stamps);
        double sum[4] = {0, 0, 0, 0};
        for (size_t c = 0; c < n_cta; ++c)
            for (int k = 0; k < 4; ++k) sum[k] += (double)(h[c * 6 + k + 1] - h[c * 6 + k]);
        // per SM: first start to last end, and the sum of its CTAs' lifetimes (2 CTAs resident -> ratio ~2 when slots never idle)
        long long t_min[256], t_max[256]; double life[256]; int cnt[256];
        for (int i = 0; i < 256; ++i) { t_min[i] = (1ll << 62); t_max[i] = 0; life[i] = 0; cnt[i] = 0; }
        for (size_t c = 0; c < n_cta; ++c) {
            const int sm = (int)(h[c * 6 + 5] & 255);
            t_min[sm] = h[c * 6] < t_min[sm] ? h[c * 6] : t_min[sm];
            t_max[sm] = h[c * 6 + 4] > t_max[sm] ? h[c * 6 + 4] : t_max[sm];
            life[sm] += (double)(h[c * 6 + 4] - h[c * 6]); ++cnt[sm];
        }
        double span = 0, lifes = 0; int n_sm = 0, cmax = 0;
        for (int i = 0; i < 256; ++i) if (cnt[i]) { span += (double)(t_max[i] - t_min[i]); lifes += life[i]; ++n_sm; cmax = cnt[i] > cmax ? cnt[i] : cmax; }
        if (FILE* f = fopen(stamp_file, "a")) {
            fprintf(f, "s2d R=%d ctas %zu on %d SMs (max %d per SM): mean clocks per CTA: data wait %.0f, pass 1 %.0f, pass 2 %.0f, conv %.0f, total %.0f | "
                       "per SM: span %.0f clk, sum of CTA lifetimes / span = %.2f\n", R, n_cta, n_sm, cmax, sum[0] / n_cta, sum[1] / n_cta, sum[2] / n_cta,
                    sum[3] / n_cta, (sum[0] + sum[1] + sum[2] + sum[3]) / n_cta, span / n_sm, lifes / span);
            fclose(f);
        }
    }
#endif
    return 0;
}

// the pyramid output is a test / debugging facility: its stores live in a second instantiation of the kernel
template <int R, int MINM = -1, int MAXM = -1, int NCONV = 0, int CIN = 0>
int launch_fast(FastParams& p, cudaStream_t stream) {
    return p.pyramid ? launch_fast_t<R, MINM, MAXM, NCONV, CIN, true>(p, stream) : launch_fast_t<R, MINM, MAXM, NCONV, CIN, false>(p, stream);
}

// canonical pool order; false if the configuration is outside the fast path (duplicate sizes within a group)
bool make_pool_spec(const int* min_sizes, int n_min, const int* max_sizes, int n_max, PoolSpec* ps) {
    memset(ps, 0, sizeof(*ps));
    ps->n_pool = n_min + n_max;
    ps->n_min = n_min;
    int k = 0;
    for (int grp = 0; grp < 2; ++grp) {
        const int* sizes = grp ? max_sizes : min_sizes;
        const int cnt = grp ? n_max : n_min, base = grp ? n_min : 0;
        for (int r = 1; r <= 16; ++r)
            for (int i = 0; i < cnt; ++i)
                if (sizes[i] / 2 == r) {
                    unsigned& mask = grp ? ps->max_mask : ps->min_mask;
                    if (mask & (1u << r)) return false;
                    mask |= 1u << r;
                    ps->radius[k] = r;
                    ps->chan[k] = base + i;
                    ++k;
                    int& rmax = grp ? ps->rmax_max : ps->rmax_min;
                    rmax = r;
                }
    }
    return k == ps->n_pool;
}

// Is this configuration on the packed fast path?  (C_in <= 2, distinct pool sizes, 0 <= slope <= 1, n_filter == 8 and a
// shared-memory footprint that lets two CTAs share an SM.)  Everything else runs the generic exact-fp32 kernel.
bool fast_path(int C_in, const int* min_sizes, int n_min, const int* max_sizes, int n_max, int n_conv, int n_filter, float slope,
               PoolSpec* ps, int* R_out) {
    if (n_filter != NF || n_conv < 1 || n_conv > MAX_CONV || n_min < 0 || n_max < 0 || n_min + n_max < 1 || n_min + n_max > MAX_POOL)
        return false;
    for (int k = 0; k < n_min + n_max; ++k) {
        const int s = k < n_min ? min_sizes[k] : max_sizes[k - n_min];
        if (s < 3 || !(s & 1) || s > 33) return false;
    }
    if (!make_pool_spec(min_sizes, n_min, max_sizes, n_max, ps)) return false;
    const int rmax = ps->rmax_min > ps->rmax_max ? ps->rmax_min : ps->rmax_max;
    const int R = rmax <= 4 ? 4 : (rmax <= 8 ? 8 : 16);
    *R_out = R;
    return C_in >= 1 && C_in <= 2 && slope >= 0.f && slope <= 1.f && make_layout(R, ps->n_pool, n_conv).total <= SMEM_MAX;
}

}  // namespace

size_t s2d_wpack_bytes(int C_in, const int* min_sizes, int n_min, const int* max_sizes, int n_max, int n_conv, int n_filter, float slope) {
    PoolSpec ps;
    int R;
    if (!min_sizes || !max_sizes || !fast_path(C_in, min_sizes, n_min, max_sizes, n_max, n_conv, n_filter, slope, &ps, &R)) return 0;
    return (size_t)w1_bytes(ps.n_pool, n_conv) + WB_BYTES;
}

int s2d_pack_weights(const float* const* w_1x1, int n_conv, const float* w_3x3, int C_in, const int* min_sizes, int n_min,
                     const int* max_sizes, int n_max, void* wpack, cudaStream_t stream) {
    KB_CHECK_ARG(w_1x1 && w_3x3 && wpack, "s2d pack: null pointer argument");
    KB_CHECK_ARG(n_conv >= 1 && n_conv <= MAX_CONV && C_in >= 1 && n_min >= 0 && n_max >= 0 && n_min + n_max >= 1 &&
                 n_min + n_max <= MAX_POOL, "s2d pack: bad configuration");
    KB_CHECK_ARG(((uintptr_t)wpack & 15) == 0, "s2d pack: wpack must be 16-byte aligned");
    PoolSpec ps;
    KB_CHECK_ARG(C_in <= 2 && make_pool_spec(min_sizes, n_min, max_sizes, n_max, &ps),
                 "s2d pack: configuration is outside the packed fast path (input_channels > 2 or duplicate pool sizes)");
    WeightSrc ws;
    memset(&ws, 0, sizeof(ws));
    for (int l = 0; l < n_conv; ++l) {
        KB_CHECK_ARG(w_1x1[l], "s2d pack: w_1x1[%d] is null", l);
        ws.w1x1[l] = w_1x1[l];
    }
    ws.w3x3 = w_3x3; ws.n_conv = n_conv; ws.n_pool = ps.n_pool; ws.C_in = C_in;
    for (int k = 0; k < MAX_POOL; ++k) ws.chan[k] = ps.chan[k];
    s2d_pack_weights_kernel<<<1, 256, 0, stream>>>(ws, (uint8_t*)wpack);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int s2d_forward(const float* x, int N, int C_in, int H, int W, const int* min_sizes, int n_min, const int* max_sizes,
                int n_max, const float* const* w_1x1, int n_conv, int n_filter, const float* w_3x3, const void* wpack, float neg_slope,
                float* out, float* pyramid, cudaStream_t stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(x && out && w_3x3 && w_1x1, "s2d: null pointer argument");
    KB_CHECK_ARG(N >= 0 && H > 0 && W > 0, "s2d: bad shape N=%d H=%d W=%d", N, H, W);
    KB_CHECK_ARG(n_filter == NF, "s2d: n_filter=%d unsupported (only %d)", n_filter, NF);
    KB_CHECK_ARG(n_conv >= 1 && n_conv <= MAX_CONV, "s2d: n_convolution=%d unsupported (1..%d)", n_conv, MAX_CONV);
    KB_CHECK_ARG(n_min >= 0 && n_max >= 0 && n_min + n_max >= 1 && n_min + n_max <= MAX_POOL,
                 "s2d: %d pool sizes unsupported (1..%d)", n_min + n_max, MAX_POOL);
    for (int k = 0; k < n_min + n_max; ++k) {
        const int s = k < n_min ? min_sizes[k] : max_sizes[k - n_min];
        KB_CHECK_ARG(s >= 3 && (s & 1) && s <= 33, "s2d: pool size %d unsupported (odd, 3..33)", s);
    }
    for (int l = 0; l < n_conv; ++l) KB_CHECK_ARG(w_1x1[l], "s2d: w_1x1[%d] is null", l);
    FastParams p;
    memset(&p, 0, sizeof(p));
    int R = 0;
    if (!fast_path(C_in, min_sizes, n_min, max_sizes, n_max, n_conv, n_filter, neg_slope, &p.ps, &R)) {
        KB_CHECK_ARG(wpack == nullptr, "s2d: packed weights given for a configuration outside the packed fast path");
        return s2d_generic_forward(x, N, C_in, H, W, min_sizes, n_min, max_sizes, n_max, w_1x1, n_conv, n_filter, w_3x3, neg_slope,
                                   out, pyramid, stream);
    }
    KB_CHECK_ARG(((uintptr_t)wpack & 15) == 0, "s2d: wpack must be 16-byte aligned");
    p.x = x; p.out = out; p.pyramid = pyramid;
    p.N = N; p.C_in = C_in; p.H = H; p.W = W;
    for (int l = 0; l < n_conv; ++l) p.ws.w1x1[l] = w_1x1[l];
    p.ws.w3x3 = w_3x3; p.ws.n_conv = n_conv; p.ws.n_pool = p.ps.n_pool; p.ws.C_in = C_in;
    for (int k = 0; k < MAX_POOL; ++k) p.ws.chan[k] = p.ps.chan[k];
    p.n_conv = n_conv; p.wpack = (const uint8_t*)wpack; p.slope = neg_slope;
    // compile-time specialisations of the shipped configurations (bash/kitti/run_kbnet_kitti_validation.sh:10-16,
    // bash/void/run_kbnet_void1500.sh:15-16, bash/void/train_kbnet_void1500.sh:21-22); everything else takes the same kernel
    // with the pool configuration read from the launch parameters
    static const bool no_spec = getenv("KBNET_S2D_NO_SPEC") != nullptr;
    const unsigned mm = p.ps.min_mask, xm = p.ps.max_mask;
    if (!no_spec && n_conv == 3) {
        if (R == 8 && mm == 0x7Cu && xm == 0x180u && C_in == 2) return launch_fast<8, 0x7C, 0x180, 3, 2>(p, stream);
        if (R == 8 && mm == 0x7Cu && xm == 0x180u && C_in == 1) return launch_fast<8, 0x7C, 0x180, 3, 1>(p, stream);
        if (R == 16 && mm == 0x180u && xm == 0x6800u && C_in == 2) return launch_fast<16, 0x180, 0x6800, 3, 2>(p, stream);
        if (R == 16 && mm == 0x380u && xm == 0x2800u && C_in == 2) return launch_fast<16, 0x380, 0x2800, 3, 2>(p, stream);
    }
    if (R == 4) return launch_fast<4>(p, stream);
    if (R == 8) return launch_fast<8>(p, stream);
    return launch_fast<16>(p, stream);
}

}  // namespace kbnet
