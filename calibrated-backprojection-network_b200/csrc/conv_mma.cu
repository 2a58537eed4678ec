// 3x3 stride-1 convolutions with FEW channels at FULL resolution (KBNet: conv0_image 3 -> 48, conv0_depth 8 -> 16,
// deconv0.conv 12 -> 12 at 352 x 1216; reference src/networks.py:364-365, src/net_utils.py:1482-1487) on the warp-level
// tensor path (mma.sync m16n8k16 bf16, fp32 accumulate, 3-term operand split).
//
// Why not the tcgen05 engine (conv_tma.cu) for these: a tcgen05.mma costs >= ~45 clocks whatever N is (profiles/
// r01_umma_rate_probe.log), so with 12-48 filters and K = 27-108 the engine is bound by the NUMBER of MMAs (3 -> 48: 249 us,
// 8 -> 16: 152 us, 12 -> 12: 154 us for 8 KITTI frames = 16-43 % of the HBM roofline).  mma.sync issues one m16n8k16 per
// 2.1 clocks per SM (profiles/r01_pipe_probe.log): these layers then sit near their HBM time.
//
// Per CTA: a 16 x 64 output tile, 256 threads.
//   1. convert   every thread loads the channels of a few pixels of the 18 x 68 halo tile straight from global memory
//                (coalesced along x, zero outside the image = the conv padding), splits them to bf16 hi + bf16 lo and stores
//                16-byte "chunk plane" entries: MODE 8: 8 channels of one pixel;  MODE 4 (C_in <= 4): plane A = 4 channels of
//                pixel p and of p + 1, plane B = 4 channels of p + 2 and zeros -- so a whole kernel row is ONE k16 step.
//                Columns are stored in an even and an odd half-row: the mma rows of a 16-pixel block are its even pixels then
//                its odd pixels, which makes the two pixels a lane owns adjacent (8-byte output stores).
//   2. mma       warp = (16-column block, row group or filter half).  For every input row the A fragments come from the planes
//                with ldmatrix (a tap is an address shift) and feed the (up to) three output rows that are open; consecutive
//                HMMAs go to different accumulators.  B fragments (weights, bf16 hi/lo, built from the fp32 OIHW tensor by
//                every CTA -- they are a few KB) live in registers.
//   3. epilogue  an output row is complete two input rows later: LeakyReLU, 8-byte NCHW stores.
#include <cuda.h>
#include <string.h>

#include "conv_common.cuh"
#include "umma.cuh"

namespace kbnet {

namespace {

using namespace ptx;

constexpr int TH = 16, TW = 64, PH = TH + 2, PW = TW + 4;   // tile, halo tile (columns x0-2 .. x0+65)
constexpr int NT = 256;
constexpr int QROW = PW;                                     // entries per plane row: 34 even columns, then 34 odd columns
constexpr int XP = PW + 4;                                   // raw (TMA) tile row pitch: the box starts at x0 - 4 (16-byte aligned coordinate)

struct MmaConvParams {
    const float* x;
    long long x_stride_n;
    const float* w;      // [F, C, 3, 3]
    float* out;
    long long out_stride_n;
    int N, C, H, W, F;
    int act;
    float slope;
    int use_tma;
    // optional operand-format copy of the output (bf16 hi / lo, [N][F/8][4 parity planes][H/2][W/2] x 16-byte entries) for a following KB block;
    // `out` may be null when only this copy is wanted
    uint8_t* of_hi;
    uint8_t* of_lo;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// (a, b) -> packed bf16 hi pair (round to nearest) and lo pair = rn(residual): 16 mantissa bits per value
__device__ __forceinline__ void split_rn(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = cvt_bf16x2(a, b);
    lo = cvt_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float leaky01(float v, float slope) { return fmaxf(v, v * slope); }

template <int MODE, int NCH>
struct Geo {
    static constexpr int NPL = MODE == 4 ? 2 : NCH;                     // chunk planes (each with a hi and a lo copy)
    static constexpr int STEPS = MODE == 4 ? 1 : (3 * NCH + 1) / 2;     // k16 steps per kernel row
    static constexpr int PLANE_BYTES = PH * QROW * 16;
};

// plane entry of halo-tile pixel (row, column c)
__device__ __forceinline__ int entry_of(int row, int c) { return row * QROW + (c & 1) * (QROW / 2) + (c >> 1); }

// MODE / NCH: operand packing (see the header);  NTW: n8 filter tiles per warp;  NSPLIT: the two warps of a column block split
// the FILTERS (each sweeps all 16 rows) instead of the rows (each sweeps 8 rows for all filters)
// OF: the instantiation that can write the operand-format copy (its address arithmetic and split stay out of the plain kernel)
template <int MODE, int NCH, int NTW, bool NSPLIT, bool OF>
__global__ void __launch_bounds__(NT, 2) conv_mma_kernel(const __grid_constant__ MmaConvParams p, const __grid_constant__ CUtensorMap tm_x) {
    using G = Geo<MODE, NCH>;
    constexpr int NPL = G::NPL, STEPS = G::STEPS;
    constexpr int NT8 = NSPLIT ? 2 * NTW : NTW;              // filter tiles of the layer
    constexpr int LR = NSPLIT ? TH : TH / 2;                 // output rows per warp
    extern __shared__ __align__(128) uint8_t smem[];
    uint4* s_hi = reinterpret_cast<uint4*>(smem);                                   // [plane][PH][QROW]
    uint4* s_lo = reinterpret_cast<uint4*>(smem + NPL * G::PLANE_BYTES);
    uint2* s_wB = reinterpret_cast<uint2*>(smem + 2 * NPL * G::PLANE_BYTES);        // [3 * STEPS][NT8][hi|lo][32 lanes]
    constexpr int WB_BYTES = 3 * STEPS * NT8 * 2 * 32 * 8;
    constexpr int CRAW = MODE == 4 ? 4 : 8 * NCH;                                   // channel planes of the raw tile
    float* s_raw = reinterpret_cast<float*>(smem + 2 * NPL * G::PLANE_BYTES + WB_BYTES);   // [C][PH][XP] fp32, filled by TMA
    const uint32_t bar = smem_u32(smem + 2 * NPL * G::PLANE_BYTES + WB_BYTES + CRAW * PH * XP * 4);

    const int tid = threadIdx.x;
    const int n = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = p.H, W = p.W, C = p.C, F = p.F;
    const size_t plane = (size_t)H * W;

    if (p.use_tma && tid == 0) {   // the raw halo tile: one 4-D box {72 columns from x0 - 4, 18 rows from y0 - 1, all C channels}, zero fill = padding
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(bar, (uint32_t)(p.C * PH * XP * 4));
        tma_load_4d(smem_u32(s_raw), &tm_x, x0 - 4, y0 - 1, 0, n, bar);
    }

    // ---- weights -> B fragments: n = lane / 4 (filter within the tile), b0 holds k-slots 2t, 2t+1 of the first half of the
    //      k16 step, b1 the same slots of the second half -----------------------------------------------------------------
    for (int i = tid; i < 3 * STEPS * NT8 * 32; i += NT) {
        const int lane = i & 31, nt = (i >> 5) % NT8, st = (i >> 5) / NT8;
        const int dy = st / STEPS, sub = st - dy * STEPS;
        const int f = 8 * nt + (lane >> 2), t = lane & 3;
        uint32_t hi[2], lo[2];
#pragma unroll
        for (int reg = 0; reg < 2; ++reg) {
            float wv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ks = 2 * t + e;
                int c = -1, dx = 0;
                if (MODE == 4) {
                    if (reg == 0) { c = ks & 3; dx = ks >> 2; }
                    else if (ks < 4) { c = ks; dx = 2; }
                } else {
                    const int item = 2 * sub + reg;
                    if (item < 3 * NCH) { dx = item / NCH; c = 8 * (item - dx * NCH) + ks; }
                }
                wv[e] = (c >= 0 && c < C && f < F) ? __ldg(p.w + (((size_t)f * C + c) * 3 + dy) * 3 + dx) : 0.f;
            }
            split_rn(wv[0], wv[1], hi[reg], lo[reg]);
        }
        s_wB[((st * NT8 + nt) * 2 + 0) * 32 + lane] = make_uint2(hi[0], hi[1]);
        s_wB[((st * NT8 + nt) * 2 + 1) * 32 + lane] = make_uint2(lo[0], lo[1]);
    }

    // ---- 1. convert the halo tile -----------------------------------------------------------------------------------
    if (p.use_tma) {
        __syncthreads();                                      // the barrier init is visible
        mbar_wait(bar, 0);
        for (int idx = tid; idx < PH * PW; idx += NT) {
            const int row = idx / PW, c = idx - row * PW;
            const int e = entry_of(row, c);
            const float* src = s_raw + row * XP + c + 2;      // raw column of image column x0 - 2 + c
            if (MODE == 4) {
                float v[3][4];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) v[d][ch] = ch < C ? src[ch * (PH * XP) + d] : 0.f;
                uint32_t h[4], l[4];
                split_rn(v[0][0], v[0][1], h[0], l[0]);
                split_rn(v[0][2], v[0][3], h[1], l[1]);
                split_rn(v[1][0], v[1][1], h[2], l[2]);
                split_rn(v[1][2], v[1][3], h[3], l[3]);
                s_hi[e] = make_uint4(h[0], h[1], h[2], h[3]);
                s_lo[e] = make_uint4(l[0], l[1], l[2], l[3]);
                split_rn(v[2][0], v[2][1], h[0], l[0]);
                split_rn(v[2][2], v[2][3], h[1], l[1]);
                s_hi[PH * QROW + e] = make_uint4(h[0], h[1], 0u, 0u);
                s_lo[PH * QROW + e] = make_uint4(l[0], l[1], 0u, 0u);
            } else {
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    float v[8];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) v[ch] = 8 * k + ch < C ? src[(8 * k + ch) * (PH * XP)] : 0.f;
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) split_rn(v[2 * q], v[2 * q + 1], h[q], l[q]);
                    s_hi[k * PH * QROW + e] = make_uint4(h[0], h[1], h[2], h[3]);
                    s_lo[k * PH * QROW + e] = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
        }
    } else {   // shapes the tensor map cannot describe (W % 4 != 0, unaligned base or strides): predicated global loads
        const float* xin = p.x + (size_t)n * p.x_stride_n;
        for (int idx = tid; idx < PH * PW; idx += NT) {
            const int row = idx / PW, c = idx - row * PW;
            const int gy = y0 - 1 + row, gx = x0 - 2 + c;
            const int e = entry_of(row, c);
            const bool rin = (unsigned)gy < (unsigned)H;
            const float* src = xin + (size_t)gy * W + gx;
            if (MODE == 4) {
                float v[3][4];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const bool in = rin && (unsigned)(gx + d) < (unsigned)W;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) v[d][ch] = (in && ch < C) ? __ldg(src + (size_t)ch * plane + d) : 0.f;
                }
                uint32_t h[4], l[4];
                split_rn(v[0][0], v[0][1], h[0], l[0]);
                split_rn(v[0][2], v[0][3], h[1], l[1]);
                split_rn(v[1][0], v[1][1], h[2], l[2]);
                split_rn(v[1][2], v[1][3], h[3], l[3]);
                s_hi[e] = make_uint4(h[0], h[1], h[2], h[3]);
                s_lo[e] = make_uint4(l[0], l[1], l[2], l[3]);
                split_rn(v[2][0], v[2][1], h[0], l[0]);
                split_rn(v[2][2], v[2][3], h[1], l[1]);
                s_hi[PH * QROW + e] = make_uint4(h[0], h[1], 0u, 0u);
                s_lo[PH * QROW + e] = make_uint4(l[0], l[1], 0u, 0u);
            } else {
                const bool in = rin && (unsigned)gx < (unsigned)W;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    float v[8];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) v[ch] = (in && 8 * k + ch < C) ? __ldg(src + (size_t)(8 * k + ch) * plane) : 0.f;
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) split_rn(v[2 * q], v[2 * q + 1], h[q], l[q]);
                    s_hi[k * PH * QROW + e] = make_uint4(h[0], h[1], h[2], h[3]);
                    s_lo[k * PH * QROW + e] = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
        }
    }
    __syncthreads();

    // ---- 2. mma + 3. epilogue ----------------------------------------------------------------------------------------
    {
        const int warp = tid >> 5, lane = tid & 31;
        const int cb = warp & 3, second = warp >> 2;
        const int row0 = NSPLIT ? 0 : LR * second;           // first output row of this warp (tile-relative)
        const int nt0 = NSPLIT ? NTW * second : 0;            // first filter tile of this warp
        const int g = lane >> 2, t = lane & 3;
        uint2 bh[3 * STEPS][NTW], bl[3 * STEPS][NTW];
#pragma unroll
        for (int st = 0; st < 3 * STEPS; ++st)
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                bh[st][j] = s_wB[((st * NT8 + nt0 + j) * 2 + 0) * 32 + lane];
                bl[st][j] = s_wB[((st * NT8 + nt0 + j) * 2 + 1) * 32 + lane];
            }
        // ldmatrix row addresses: matrix mi = lane / 8 -> row half (mi & 1) = even / odd pixels of the block, k half (mi >> 1)
        const int mi = lane >> 3, mr = lane & 7;
        uint32_t a_off[STEPS];                                // byte offset of this lane's row inside the hi (or lo) plane set
#pragma unroll
        for (int sub = 0; sub < STEPS; ++sub) {
            int pl, dx;
            if (MODE == 4) { pl = mi >> 1; dx = 0; }
            else {
                int item = 2 * sub + (mi >> 1);
                if (item >= 3 * NCH) item = 3 * NCH - 1;       // the unused half of an odd last step: weights are zero, address stays valid
                dx = item / NCH; pl = item - dx * NCH;
            }
            const int c = 16 * cb + 2 * mr + (mi & 1) + 1 + dx;
            a_off[sub] = (uint32_t)(pl * PH * QROW + entry_of(row0, c)) * 16u;
        }
        const uint32_t hi_base = smem_u32(s_hi), lo_base = smem_u32(s_lo);
        const int gxa = x0 + 16 * cb + 2 * g;                 // the lane's two output pixels: gxa, gxa + 1
        float* optr[NTW];                                     // filter 8 (nt0 + j) + 2t at this lane's first pixel, advanced row by row
#pragma unroll
        for (int j = 0; j < NTW; ++j)
            optr[j] = p.out + (size_t)n * p.out_stride_n + (size_t)(8 * (nt0 + j) + 2 * t) * plane + (size_t)(y0 + row0) * W + gxa;
        const bool full = y0 + TH <= H && x0 + TW <= W && (W & 1) == 0 && ((uintptr_t)p.out & 7) == 0 && (p.out_stride_n & 1) == 0;
        const bool want_nchw = !OF || p.out != nullptr, want_of = OF && p.of_hi != nullptr;
        const int of_cg = (F + 7) >> 3;
        // operand-format entry of (chunk, parity plane, y / 2, x / 2): 16 bytes; this lane owns bytes [4t, 4t + 4) of the entries of
        // its two pixels, which sit in the even-column and the odd-column plane of the row's parity (H and W are even)
        const size_t of_plane = (size_t)(H >> 1) * (W >> 1) * 16;
        const size_t of_base = want_of ? ((size_t)n * of_cg + nt0) * 4 * of_plane + (size_t)(gxa >> 1) * 16 + 4 * t : 0;
        const bool in_a = gxa < W, in_b = gxa + 1 < W;
        float acc[3][NTW][4];                                 // the three open output rows
#pragma unroll
        for (int qr = 0; qr < LR + 2; ++qr) {
            if (qr < LR) {
#pragma unroll
                for (int j = 0; j < NTW; ++j) acc[qr % 3][j][0] = acc[qr % 3][j][1] = acc[qr % 3][j][2] = acc[qr % 3][j][3] = 0.f;
            }
#pragma unroll
            for (int sub = 0; sub < STEPS; ++sub) {
                uint32_t ah[4], al[4];
                ldsm_x4(hi_base + a_off[sub] + (uint32_t)(qr * QROW * 16), ah);
                ldsm_x4(lo_base + a_off[sub] + (uint32_t)(qr * QROW * 16), al);
#pragma unroll
                for (int term = 0; term < 3; ++term) {        // a_hi b_hi, a_lo b_hi, a_hi b_lo: walk the accumulators round-robin
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int li = qr - dy;
                        if (li >= 0 && li < LR) {
#pragma unroll
                            for (int j = 0; j < NTW; ++j) {
                                const uint2 b = term == 2 ? bl[dy * STEPS + sub][j] : bh[dy * STEPS + sub][j];
                                hmma(acc[li % 3][j], term == 1 ? al : ah, b.x, b.y);
                            }
                        }
                    }
                }
            }
            if (qr >= 2) {                                    // output row qr - 2 is complete
                const int li = qr - 2;
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const int f0 = 8 * (nt0 + j) + 2 * t;     // c0, c2 -> filter f0 at pixels gxa, gxa + 1;  c1, c3 -> filter f0 + 1
                    float v0 = acc[li % 3][j][0], v1 = acc[li % 3][j][1], v2 = acc[li % 3][j][2], v3 = acc[li % 3][j][3];
                    if (p.act) { v0 = leaky01(v0, p.slope); v1 = leaky01(v1, p.slope); v2 = leaky01(v2, p.slope); v3 = leaky01(v3, p.slope); }
                    float* d0 = optr[j];
                    float* d1 = d0 + plane;
                    optr[j] += W;
                    if (OF && want_of && y0 + row0 + li < H && f0 < F) {
                        uint32_t h0, l0, h1, l1;
                        split_rn(v0, v1, h0, l0);             // pixel gxa: filters f0, f0 + 1
                        split_rn(v2, v3, h1, l1);             // pixel gxa + 1
                        const int y = y0 + row0 + li;
                        const size_t e = of_base + ((size_t)j * 4 + (size_t)(y & 1) * 2) * of_plane + (size_t)(y >> 1) * (size_t)(W >> 1) * 16;
                        if (in_a) { *reinterpret_cast<uint32_t*>(p.of_hi + e) = h0; *reinterpret_cast<uint32_t*>(p.of_lo + e) = l0; }
                        if (in_b) { *reinterpret_cast<uint32_t*>(p.of_hi + e + of_plane) = h1; *reinterpret_cast<uint32_t*>(p.of_lo + e + of_plane) = l1; }
                    }
                    if (OF && !want_nchw) continue;
                    if (full && f0 + 1 < F) {
                        *reinterpret_cast<float2*>(d0) = make_float2(v0, v2);
                        *reinterpret_cast<float2*>(d1) = make_float2(v1, v3);
                    } else if (y0 + row0 + li < H) {
                        if (f0 < F) { if (in_a) d0[0] = v0; if (in_b) d0[1] = v2; }
                        if (f0 + 1 < F) { if (in_a) d1[0] = v1; if (in_b) d1[1] = v3; }
                    }
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    return fn;
}

// 4-D tiled map over x [N][C][H][W] fp32 (sample pitch x_stride_n), box = {XP, PH, C, 1}; out-of-bounds elements read as zero
bool make_map(CUtensorMap* tm, const MmaConvParams& p) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    if ((p.W & 3) || ((uintptr_t)p.x & 15) || (p.x_stride_n & 3)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.C, (cuuint64_t)p.N};
    const cuuint64_t strides[3] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.W * p.H * 4, (cuuint64_t)p.x_stride_n * 4};
    const cuuint32_t box[4] = {(cuuint32_t)XP, (cuuint32_t)PH, (cuuint32_t)p.C, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE, int NCH, int NTW, bool NSPLIT>
int launch(MmaConvParams& p, cudaStream_t stream) {
    using G = Geo<MODE, NCH>;
    constexpr int NT8 = NSPLIT ? 2 * NTW : NTW;
    constexpr int CRAW = MODE == 4 ? 4 : 8 * NCH;
    const int smem = 2 * G::NPL * G::PLANE_BYTES + 3 * G::STEPS * NT8 * 2 * 32 * 8 + CRAW * PH * XP * 4 + 16;
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    p.use_tma = make_map(&tm, p) ? 1 : 0;
    auto kern = p.of_hi ? conv_mma_kernel<MODE, NCH, NTW, NSPLIT, true> : conv_mma_kernel<MODE, NCH, NTW, NSPLIT, false>;
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int tiles_x = cdiv(p.W, TW), tiles_y = cdiv(p.H, TH);
    KB_CHECK_ARG(tiles_y <= 65535 && p.N <= 65535, "conv_mma: more than 65535 tile rows or samples in one launch");
    kern<<<dim3((unsigned)tiles_x, (unsigned)tiles_y, (unsigned)p.N), NT, smem, stream>>>(p, tm);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace

// few channels in, few filters out, 3x3 stride 1, one source at the conv resolution
bool conv_mma_supported(int C, int F, int ksize, int stride) {
    if (ksize != 3 || stride != 1 || C < 1 || F < 1) return false;
    // C in 9..16 (two chunk planes, 9 k16 steps per output row) is implemented but measured slower than the tcgen05 engine
    // (12 -> 12 at 352 x 1216 x 8: 174 us vs 154 us), so it is not offered
    return (C <= 4 && F <= 48) || (C <= 8 && F <= 16);
}

int conv_mma_launch(const float* x, long long x_stride_n, int N, int C, int H, int W, const float* w, int F, int act, float slope, float* out,
                    long long out_stride_n, cudaStream_t stream, void* of_hi, void* of_lo) {
    KB_CHECK_ARG(conv_mma_supported(C, F, 3, 1), "conv_mma: C=%d F=%d unsupported", C, F);
    KB_CHECK_ARG((of_hi != nullptr) == (of_lo != nullptr) &&
                 (of_hi == nullptr || ((F & 7) == 0 && (((uintptr_t)of_hi | (uintptr_t)of_lo) & 15) == 0 && !(H & 1) && !(W & 1))),
                 "conv_mma: the operand-format output needs both halves, 16-byte alignment, F %% 8 == 0 and even H, W");
    KB_CHECK_ARG(out != nullptr || of_hi != nullptr, "conv_mma: no output");
    MmaConvParams p;
    p.x = x; p.x_stride_n = x_stride_n; p.w = w; p.out = out; p.out_stride_n = out_stride_n;
    p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.act = act; p.slope = slope;
    p.of_hi = (uint8_t*)of_hi; p.of_lo = (uint8_t*)of_lo;
    if (C <= 4) {
        if (F <= 16) return launch<4, 1, 1, true>(p, stream);
        if (F <= 32) return launch<4, 1, 2, true>(p, stream);
        return launch<4, 1, 3, true>(p, stream);
    }
    return F <= 8 ? launch<8, 1, 1, false>(p, stream) : launch<8, 1, 2, false>(p, stream);
}

}  // namespace kbnet
