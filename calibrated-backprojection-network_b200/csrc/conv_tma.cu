// Tensor-core convolution engine, second generation ("v2"): persistent, TMA-fed, warp-specialised, table-driven.
//
//   out_j[n, f, oy, ox] = act( sum_{c, ky, kx} w_j[f, c, ky, kx] * in_j[n, c, oy*s - pad + ky, ox*s - pad + kx] )     j = 0 .. n_convs-1
//
// One launch runs a small PROGRAM: a list of K16 stages (two 8-channel chunks each) and, per stage, the convolutions
// ("uses") that consume it.  A plain convolution is a program with one use per stage; the Calibrated Backprojection block
// (reference src/net_utils.py:1343-1371) is ONE program in which the image tile is staged once and feeds both conv_image
// (3x3 stride 2, nine taps) and conv_fused (1x1 stride 2 = the centre tap of the same parity planes), the depth tile
// feeds conv_depth and -- on the CUDA cores of the converter warps -- z = leaky(w_proj . depth), the camera coordinates
// are generated from K^-1 in registers, xyz = coordinates * z goes straight into an operand plane, and three accumulators
// live side by side in TMEM.  No intermediate (coordinates, z, xyz, any torch.cat) ever exists in global memory.
//
// Data path per CTA (one CTA per SM, looping over tiles):
//   HBM --TMA (cp.async.bulk.tensor.4d, fp32 NCHW boxes, zero fill = conv padding)--> raw ring in shared memory
//       --converter warps (fp32 -> bf16 hi + bf16 lo, transposed to K-major "chunk planes")--> A ring
//       --tcgen05.mma (one thread; 3x3 taps are row-shifted descriptors of the same plane, no im2col)--> TMEM accumulators
//       --epilogue warps (tcgen05.ld, LeakyReLU, coalesced NCHW stores)--> HBM,
//   with the accumulator double-buffered in TMEM so that the epilogue of tile i overlaps the main loop of tile i+1, and
//   the weights streamed as pre-packed bf16 hi/lo tiles through their own bulk-copy ring.
//
// Arithmetic: a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with fp32 accumulation (16 mantissa bits per operand, ~1e-5
// relative).  When the filter count is small the three products cost two MMAs: B' = [b_hi | b_lo] doubles N
// (free below N ~ 48, profiles/r01_umma_probe.md) and the epilogue adds the two column groups ("dual" mode).
//
// Roles (512 threads): warps 0-7 converters | warp 8 activation TMA | warp 9 weight TMA | warp 10 MMA issuer + TMEM
// allocator | warp 11 operand-format TMA (sources already stored as bf16 hi/lo operand planes, operand_format.cu) |
// warps 12-15 epilogue (TMEM lane quarter = warp % 4).  (warp counts: conv_warps(), cta_threads())
// The MMA issuer walks a command table the host flattens from the program for the chosen tile plan (build_commands).
//
// What bounds the low-N / low-K programs (KB blocks, measured on B200, profiles/r02b_kb_knockout.txt): not one role but the
// hand-over skeleton.  KB block 2 takes 196 us; with the MMAs compiled out 158 us, with the conversion compiled out 171 us,
// with the epilogue stores compiled out 190 us, and with ALL THREE compiled out (only TMA loads, weight streaming and mbarrier
// hand-overs left) still 117 us: every tile of 124 output pixels streams 240 KB of packed weights through a 2 x 24 KB ring
// (22 refills of ~1.2 k clk latency, two in flight) and crosses 9 operand stages of a 2-deep ring, each hand-over costing a
// ~250 clk mbarrier hop.  Shared memory (227 KB) leaves no room for deeper rings at this tile size, TMEM (3 accumulators = 224
// columns) none for a second M-block per tile.  Compile-time switches for these experiments: KBNET_V2_KO_MMA / _KO_CONV / _KO_EPI
// (results are garbage with them), KBNET_V2_WAIT_* (wait flavours), KBNET_V2_STAMPS (role timelines).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "conv_common.cuh"
#include "umma.cuh"

namespace kbnet {

using namespace ptx;

namespace {

constexpr int V2_MAX_STAGES = 72;
constexpr int V2_MAX_CONVS = 4;
constexpr int V2_MAX_USES = 4;
constexpr int V2_MAX_MAPS = 8;   // 0-3: fp32 NCHW sources (4-D maps); 4-7: operand-format sources (5-D maps: hi, lo per tensor)
// converter warps (a multiple of 4: the epilogue warps sit at conv_warps + 4 .. + 7 and address TMEM lane quarter warp % 4).
// 12 warps were measured for the stride-2 programs (they fit: ~80 registers) and changed nothing -- the KB block is bound by the
// latency chain raw TMA -> convert -> MMA -> commit of its two operand stages, not by converter throughput.
__host__ __device__ constexpr int conv_warps(int) { return 8; }
__host__ __device__ constexpr int cta_threads(int geo) { return (conv_warps(geo) + 8) * 32; }
constexpr int V2_MAX_RAW = 6, V2_MAX_A = 3, V2_MAX_W = 8;
constexpr int V2_HEADER = 2048;
constexpr int V2_SMEM_LIMIT = 227 * 1024;
constexpr int V2_MAX_ZC = 256;  // channels of the z projection

enum { G_S1K3 = 0, G_S1K1 = 1, G_S2K3 = 2, G_S2K1 = 3 };
enum { CK_NONE = 0, CK_TMA = 1, CK_TMA_P0 = 2, CK_COORDS = 3, CK_XYZ = 4, CK_XYZ_T = 5,
       CK_OF = 6, CK_OF_P0 = 7 };   // operand-format sources: TMA straight into the operand planes (all four parity planes / plane 0 only)
constexpr uint16_t TAPS_ALL9 = 0x1FF, TAPS_ONE = 0x001, TAPS_CENTER = 0x010;  // bit t = tap (ky, kx) = (t / 3, t % 3) of the launch geometry

struct V2Chunk {
    uint8_t kind, tmap, nch, cls;  // cls: 0 = source at the conv input resolution, 1 = source at half resolution (nearest 2x upsample)
    uint16_t c0;
    uint8_t zacc, pad;
};
struct V2Use {
    uint8_t conv, first;
    uint16_t tapmask;  // which taps of the geometry this convolution reads from the stage; weight tiles follow in bit order
    uint32_t wtile;    // index of the first weight tile of this (conv, stage) in the conv's packed buffer
};
struct V2Stage {
    V2Chunk ch[2];
    V2Use use[V2_MAX_USES];
    int n_use, pad;
};
struct V2Conv {
    const uint8_t* wpack;  // [tile][g = f/8][hi|lo][chunk 0|1][8 rows][16 B]   (conv_pack_weights)
    float* out;
    long long out_stride_n;
    int F, Npad, Nsl;  // Nsl = filters per split part
    int col;           // first TMEM column of this conv inside the column group of one M-block
    int act, ntaps;
    // output addressing: out + n * out_stride_n + f * oplane + o_off + oy * o_row_pitch + ox * o_col_mul   (oy, ox on the launch grid)
    long long oplane, o_off;
    int o_row_pitch, o_col_mul;
    // optional operand-format copy of the output for the next layer (operand_format.cu): 16-byte entries (8 filters, bf16) in a hi
    // and a lo tensor, [N][of_cg][4 parity planes][Hout/2][Wout/2], or with of_even only the even pixels [N][of_cg][Hout/2][Wout/2];
    // `out` may be null when only this copy is wanted
    uint8_t* of_hi;
    uint8_t* of_lo;
    int of_cg, of_even;
    int T, pad_t;      // taps of this convolution per weight-ring slot (a slot holds p.w_bytes: convolutions with fewer filters pack more taps)
};
// One entry of the MMA issuer's command table (built on the host for the chosen tile plan, read from the kernel parameters):
// one tap of one convolution for all M-blocks of the tile.  The issuing thread does no bookkeeping beyond what the flags ask for.
//   x = a_rel | flags << 16   a_rel: operand-plane offset of the tap (16-byte rows, relative to the A stage)
//   y = b_rel | dcol << 16    b_rel: offset of the tap's weights inside the weight-ring slot (16-byte units); dcol: TMEM column
//   z = instruction descriptor (N of this convolution's filter-split part; 2N in dual mode)
enum { CMD_ACCUM = 1, CMD_WAIT_A = 2, CMD_COMMIT_A = 4, CMD_WAIT_W = 8, CMD_COMMIT_W = 16 };
constexpr int V2_MAX_CMDS = 1152;

struct V2Params {
    uint4 cmd[V2_MAX_CMDS + 1];   // + one padding entry: the issuer prefetches command i + 1
    int n_cmds, pad_cmd[3];
    V2Stage stage[V2_MAX_STAGES];
    V2Conv conv[V2_MAX_CONVS];
    int n_stages, n_convs, geo, dual;
    int N, Hin, Win, Hout, Wout;
    int R, TW, P, MB, n_entries, plane_px, n_planes;
    int tiles_x, tiles_y, n_split;
    long long n_tiles;
    int BW[2], BH[2];
    int raw_bytes, a_bytes, w_bytes;  // bytes per ring slot
    int n_raw, n_a, n_w;
    int cols_block, acc_bufs, tmem_cols;
    int T;  // taps per weight-ring slot for the 9-tap uses (1, 3 or 9); single-tap uses take one slot each
    float slope;
    const float* kinv;    // [N, 9]   (CK_COORDS / CK_XYZ)
    const float* w_proj;  // [C_z]    (zacc)
    int C_z;
    int has_of;           // some chunk is CK_OF / CK_OF_P0: warp CW + 3 runs the operand-format producer and arrives on every A_FULL
    int of_box_bytes;     // bytes of one operand-format box = P * (R + 1) * 16
    long long* dbg;       // optional timeline buffer (KBNET_V2_TIMELINE): [role 0..4][4096] clock64 stamps of CTA 0
};
// Weight-ring slot formation, shared by the weight producer (device) and build_commands (host): the taps of a use go into slots
// of cv.T taps; the LAST slot of a use also takes the following uses of the same stage as long as each of them is a single
// slot by itself and the bytes fit -- one mbarrier hand-over instead of several (a hop costs the MMA issuer ~250 clk).
// Returns the number of following uses merged into the slot that starts at tap t0 of use u, and the slot's total bytes.
__host__ __device__ inline int w_slot_merge(const V2Params& p, const V2Stage& st, int u, int t0, int nt, int ntap, uint32_t* total_bytes) {
    uint32_t total = (uint32_t)nt * (uint32_t)p.conv[st.use[u].conv].Nsl * 64u;
    int n_merge = 0;
    if (t0 + nt >= ntap) {
        for (int v = u + 1; v < st.n_use; ++v) {
            const V2Conv& c2 = p.conv[st.use[v].conv];
            unsigned m = st.use[v].tapmask;
            int nt2 = 0;
            for (; m; m &= m - 1) ++nt2;
            const uint32_t b2 = (uint32_t)nt2 * (uint32_t)c2.Nsl * 64u;
            if (nt2 > c2.T || total + b2 > (uint32_t)p.w_bytes) break;
            total += b2;
            ++n_merge;
        }
    }
    *total_bytes = total;
    return n_merge;
}

struct alignas(64) V2Maps {
    CUtensorMap m[V2_MAX_MAPS];
};

// barrier slots
constexpr int B_RAW_FULL = 0, B_RAW_EMPTY = B_RAW_FULL + V2_MAX_RAW, B_A_FULL = B_RAW_EMPTY + V2_MAX_RAW, B_A_EMPTY = B_A_FULL + V2_MAX_A,
              B_W_FULL = B_A_EMPTY + V2_MAX_A, B_W_EMPTY = B_W_FULL + V2_MAX_W, B_ACC_FULL = B_W_EMPTY + V2_MAX_W, B_ACC_EMPTY = B_ACC_FULL + 2,
              B_COUNT = B_ACC_EMPTY + 2;
static_assert(B_COUNT * 8 <= 512, "barrier area");

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void store_split8(uint8_t* hi_dst, uint8_t* lo_dst, const float (&v)[8]) {
    const uint2 a = split_bf16x2(v[0], v[1]), b = split_bf16x2(v[2], v[3]);
    const uint2 c = split_bf16x2(v[4], v[5]), d = split_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(a.x, b.x, c.x, d.x);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(a.y, b.y, c.y, d.y);
}

// exactly one lane of the (converged) warp gets true; lets the compiler keep warp-uniform operands in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}

// ring position kept incrementally (a runtime modulo costs ~30 dependent instructions on the single issuing warp)
struct Ring {
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ void advance(uint32_t n) {
        if (++slot == n) { slot = 0; phase ^= 1u; }
    }
};

// clock64 stamps of CTA 0 for tools/v2_timeline.py: compiled in only with -DKBNET_V2_STAMPS (they sit in the hot loops of every role)
#ifdef KBNET_V2_STAMPS
#define V2_STAMP(role, idx)                                                                             \
    do {                                                                                                  \
        if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 4096) p.dbg[(role) * 4096 + (idx)] = clock64(); \
    } while (0)
#else
#define V2_STAMP(role, idx) \
    do {                    \
    } while (0)
#endif

// wait flavours of the roles (see umma.cuh).  Measured on B200 (KB block 2, 196 us with plain spinning everywhere): backing off in
// the producers / epilogue or polling with one lane per converter warp all made the block SLOWER (324 us with all three) -- every
// role is on some latency chain -- so plain spinning is the default; the variants stay selectable at compile time.
#ifdef KBNET_V2_WAIT_CONV1
#define V2_WAIT_CONV(bar, par) mbar_wait_warp(bar, par, lane)
#else
#define V2_WAIT_CONV(bar, par) mbar_wait(bar, par)
#endif
#ifdef KBNET_V2_WAIT_PROD
#define V2_WAIT_RELAXED(bar, par, ns) mbar_wait_relaxed(bar, par, ns)
#else
#define V2_WAIT_RELAXED(bar, par, ns) mbar_wait(bar, par)
#endif
#ifdef KBNET_V2_WAIT_EPI
#define V2_WAIT_EPI(bar, par)                      \
    do {                                           \
        if (lane == 0) mbar_wait_relaxed(bar, par, 128); \
        __syncwarp();                              \
    } while (0)
#else
#define V2_WAIT_EPI(bar, par) mbar_wait(bar, par)
#endif

struct TileCoord {
    int n, oy0, ox0, split;
};
__device__ __forceinline__ TileCoord decode_tile(const V2Params& p, long long tile) {
    TileCoord t;
    t.split = (int)(tile % p.n_split);
    long long r = tile / p.n_split;
    const int tx = (int)(r % p.tiles_x);
    r /= p.tiles_x;
    const int ty = (int)(r % p.tiles_y);
    t.n = (int)(r / p.tiles_y);
    t.oy0 = ty * p.R;
    t.ox0 = tx * p.TW;
    return t;
}

// OF: the instantiation with the operand-format producer role and epilogue copy (stride-2 3x3 programs only); the plain kernels
// carry none of that code -- their hot loops are sensitive to every KB of instructions (see the header)
template <int GEO, bool DUAL, bool OF = false>
__global__ void __launch_bounds__(cta_threads(GEO), 1) conv_v2_kernel(const __grid_constant__ V2Params p, const __grid_constant__ V2Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t bar0 = smem_u32(smem);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    float* s_wz = reinterpret_cast<float*>(smem + 1024);  // z-projection weights, zero padded
    uint8_t* sRaw = smem + V2_HEADER;
    uint8_t* sA = sRaw + (size_t)p.n_raw * p.raw_bytes;
    uint8_t* sW = sA + (size_t)p.n_a * p.a_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int CW = conv_warps(GEO), CONV_THREADS = CW * 32;  // roles: warps 0..CW-1 converters, CW raw TMA, CW+1 weights, CW+2 MMA, CW+4..CW+7 epilogue

    if (tid == 0) {
        for (int i = 0; i < V2_MAX_RAW; ++i) { mbar_init(BAR(B_RAW_FULL + i), 1); mbar_init(BAR(B_RAW_EMPTY + i), CW); }
        for (int i = 0; i < V2_MAX_A; ++i) { mbar_init(BAR(B_A_FULL + i), CW + (OF && p.has_of ? 1 : 0)); mbar_init(BAR(B_A_EMPTY + i), 1); }
        for (int i = 0; i < V2_MAX_W; ++i) { mbar_init(BAR(B_W_FULL + i), 1); mbar_init(BAR(B_W_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_ACC_FULL + i), 1); mbar_init(BAR(B_ACC_EMPTY + i), 4); }
        fence_mbar_init();
    }
    if (tid < V2_MAX_ZC) s_wz[tid] = (p.w_proj != nullptr && tid < p.C_z) ? __ldg(p.w_proj + tid) : 0.f;
    if (warp == CW + 2) {
        tmem_alloc_dyn(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long tile0 = blockIdx.x, tstep = gridDim.x;

    if (warp < CW) {
        // =================================================================================================================
        // converters: raw fp32 tile (shared) -> bf16 hi/lo chunk planes (shared), plus the generated chunks and z
        // =================================================================================================================
        constexpr int MAXI = GEO == G_S2K3 ? 1 : (GEO == G_S2K1 ? 2 : 4);  // entries per converter thread (the four-plane path is the largest
                                                                            // piece of code in the kernel: one item keeps it out of the I-cache's way)
        int it_e[MAXI], it_ei[MAXI], it_ej[MAXI], ro0[MAXI], ro1[MAXI];
        int n_it = 0;
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
            const int e = tid + i * CONV_THREADS;
            it_e[i] = e;
            const int ei = e / p.P, ej = e - ei * p.P;
            it_ei[i] = ei; it_ej[i] = ej;
            if (GEO == G_S1K3) { ro0[i] = ei * p.BW[0] + ej + 3; ro1[i] = ((ei + 1) >> 1) * p.BW[1] + ((ej + 1) >> 1) + 3; }
            else if (GEO == G_S1K1) { ro0[i] = ei * p.BW[0] + ej; ro1[i] = 0; }
            else if (GEO == G_S2K3) { ro0[i] = (2 * ei) * p.BW[0] + 2 * ej + 2; ro1[i] = ei * p.BW[1] + 2 * ej + 2; }
            else { ro0[i] = (2 * ei) * p.BW[0] + 2 * ej; ro1[i] = 0; }
            if (e < p.n_entries) n_it = i + 1;
        }
        const int cp0 = p.BW[0] * p.BH[0], cp1 = p.BW[1] * p.BH[1];  // channel pitch of the raw tile (floats)
        const size_t plane_b = (size_t)p.plane_px * 16;
        Ring raw_r, a_r;
        int dbg_i = 0;
        const bool dbg_on = tid == 0;
        for (long long tile = tile0; tile < p.n_tiles; tile += tstep) {
            const TileCoord tc = decode_tile(p, tile);
            float zacc[MAXI];
#pragma unroll
            for (int i = 0; i < MAXI; ++i) zacc[i] = 0.f;
            float kv[9];
            if (p.kinv != nullptr) {
#pragma unroll
                for (int i = 0; i < 9; ++i) kv[i] = __ldg(p.kinv + (size_t)tc.n * 9 + i);
            }
            for (int s = 0; s < p.n_stages; ++s) {
                const V2Stage& st = p.stage[s];
                const int aslot = (int)a_r.slot;
                V2_WAIT_CONV(BAR(B_A_EMPTY + aslot), a_r.phase ^ 1u);
                if (dbg_on) { V2_STAMP(0, dbg_i); ++dbg_i; }
                uint8_t* A = sA + (size_t)aslot * p.a_bytes;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const V2Chunk ch = st.ch[h];
                    const int kind = ch.kind;
                    if (OF && (kind == CK_OF || kind == CK_OF_P0)) continue;   // lands in the operand planes by TMA (warp CW + 3)
                    const bool is_tma = kind == CK_TMA || kind == CK_TMA_P0 || kind == CK_XYZ_T;
#ifdef KBNET_V2_KO_CONV
                    if (is_tma) {
                        const int rs = (int)raw_r.slot;
                        V2_WAIT_CONV(BAR(B_RAW_FULL + rs), raw_r.phase);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(B_RAW_EMPTY + rs));
                        raw_r.advance((uint32_t)p.n_raw);
                    }
                    continue;
#endif
                    const float* raw = nullptr;
                    int rslot = 0;
                    if (is_tma) {
                        rslot = (int)raw_r.slot;
                        V2_WAIT_CONV(BAR(B_RAW_FULL + rslot), raw_r.phase);
                        if (dbg_on) { V2_STAMP(0, dbg_i); ++dbg_i; }
                        raw = reinterpret_cast<const float*>(sRaw + (size_t)rslot * p.raw_bytes);
                    }
                    // destination of (plane pl, this chunk, entry e): hi at A + ((pl*4 + h) * plane_px + e) * 16, lo two planes further
                    if (GEO != G_S2K3) {
                        // single-plane geometries: ALL loads of this thread's items first (the compiler cannot hoist them above the
                        // operand-plane stores itself: both live in shared memory), then convert and store
                        float v[MAXI][8];
#pragma unroll
                        for (int i = 0; i < MAXI; ++i) {
                            if (i < n_it && kind == CK_TMA) {
                                const float* r = raw + (ch.cls ? ro1[i] : ro0[i]);
                                const int cp = ch.cls ? cp1 : cp0;
#pragma unroll
                                for (int c = 0; c < 8; ++c) v[i][c] = r[c * cp];
                            } else {
#pragma unroll
                                for (int c = 0; c < 8; ++c) v[i][c] = 0.f;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < MAXI; ++i) {
                            if (i < n_it) {
                                uint8_t* d = A + ((size_t)h * p.plane_px + it_e[i]) * 16;
                                store_split8(d, d + 2 * plane_b, v[i]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < MAXI; ++i) {
                        if (GEO == G_S2K3 && i < n_it && it_e[i] < p.n_entries) {
                            const int e = it_e[i];
                            {  // G_S2K3: four parity planes
                                const int y0 = 2 * (tc.oy0 - 1 + it_ei[i]), x0 = 2 * (tc.ox0 - 1 + it_ej[i]);
                                if (kind == CK_TMA) {
                                    const float* r = raw + ro0[i];
                                    float2 va[8], vb[8];
#pragma unroll
                                    for (int c = 0; c < 8; ++c) {
                                        va[c] = *reinterpret_cast<const float2*>(r + c * cp0);
                                        vb[c] = *reinterpret_cast<const float2*>(r + c * cp0 + p.BW[0]);
                                    }
                                    float v[8];
#pragma unroll
                                    for (int pl = 0; pl < 4; ++pl) {
#pragma unroll
                                        for (int c = 0; c < 8; ++c) v[c] = pl == 0 ? va[c].x : (pl == 1 ? va[c].y : (pl == 2 ? vb[c].x : vb[c].y));
                                        if (pl == 0 && ch.zacc) {
                                            float a = zacc[i];
#pragma unroll
                                            for (int c = 0; c < 8; ++c) a = fmaf(s_wz[ch.c0 + c], v[c], a);
                                            zacc[i] = a;
                                        }
                                        uint8_t* d = A + ((size_t)(pl * 4 + h) * p.plane_px + e) * 16;
                                        store_split8(d, d + 2 * plane_b, v);
                                    }
                                } else if (kind == CK_COORDS) {
#pragma unroll
                                    for (int pl = 0; pl < 4; ++pl) {
                                        const int y = y0 + (pl >> 1), x = x0 + (pl & 1);
                                        const bool in = (unsigned)y < (unsigned)p.Hin && (unsigned)x < (unsigned)p.Win;
                                        float v[8];
#pragma unroll
                                        for (int c = 0; c < 8; ++c) v[c] = 0.f;
                                        if (in) {
#pragma unroll
                                            for (int c = 0; c < 3; ++c) v[c] = fmaf(kv[3 * c], (float)x, fmaf(kv[3 * c + 1], (float)y, kv[3 * c + 2]));
                                        }
                                        uint8_t* d = A + ((size_t)(pl * 4 + h) * p.plane_px + e) * 16;
                                        store_split8(d, d + 2 * plane_b, v);
                                    }
                                } else {  // plane 0 only: CK_TMA_P0, CK_XYZ, CK_XYZ_T, CK_NONE (zeros)
                                    float v[8];
#pragma unroll
                                    for (int c = 0; c < 8; ++c) v[c] = 0.f;
                                    if (kind == CK_TMA_P0) {
                                        const float* r = raw + (ch.cls ? ro1[i] : ro0[i]);
                                        const int cp = ch.cls ? cp1 : cp0;
#pragma unroll
                                        for (int c = 0; c < 8; ++c) v[c] = r[c * cp];
                                    } else if (kind == CK_XYZ || kind == CK_XYZ_T) {
                                        const float z = leaky(zacc[i], p.slope);
                                        const bool in = (unsigned)y0 < (unsigned)p.Hin && (unsigned)x0 < (unsigned)p.Win;
                                        if (kind == CK_XYZ) {
                                            if (in) {
#pragma unroll
                                                for (int c = 0; c < 3; ++c)
                                                    v[c] = fmaf(kv[3 * c], (float)x0, fmaf(kv[3 * c + 1], (float)y0, kv[3 * c + 2])) * z;
                                            }
                                        } else {
                                            const float* r = raw + ro0[i];
#pragma unroll
                                            for (int c = 0; c < 3; ++c) v[c] = r[c * cp0] * z;
                                        }
                                    }
                                    uint8_t* d = A + ((size_t)h * p.plane_px + e) * 16;
                                    store_split8(d, d + 2 * plane_b, v);
                                    if (kind == CK_NONE) {  // a zero chunk is read by every tap of the stage: clear all planes
                                        const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                                        for (int pl = 1; pl < 4; ++pl) {
                                            uint8_t* dz = A + ((size_t)(pl * 4 + h) * p.plane_px + e) * 16;
                                            *reinterpret_cast<uint4*>(dz) = zz;
                                            *reinterpret_cast<uint4*>(dz + 2 * plane_b) = zz;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    if (is_tma) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(B_RAW_EMPTY + rslot));
                        raw_r.advance((uint32_t)p.n_raw);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_A_FULL + aslot));
                if (dbg_on) { V2_STAMP(0, dbg_i); ++dbg_i; }
                a_r.advance((uint32_t)p.n_a);
            }
        }
    } else if (warp == CW) {
        // =================================================================================================================
        // activation TMA producer (one thread)
        // =================================================================================================================
        if (elect_one()) {
            Ring raw_r;
            int dbg_i = 0;
            const uint32_t bytes0 = (uint32_t)(p.BW[0] * p.BH[0] * 32), bytes1 = (uint32_t)(p.BW[1] * p.BH[1] * 32);
            for (long long tile = tile0; tile < p.n_tiles; tile += tstep) {
                const TileCoord tc = decode_tile(p, tile);
                int x0c, y0c;
                if (GEO == G_S1K3) { x0c = tc.ox0 - 4; y0c = tc.oy0 - 1; }
                else if (GEO == G_S1K1) { x0c = tc.ox0; y0c = tc.oy0; }
                else if (GEO == G_S2K3) { x0c = 2 * tc.ox0 - 4; y0c = 2 * tc.oy0 - 2; }
                else { x0c = 2 * tc.ox0; y0c = 2 * tc.oy0; }
                // class 1: half-resolution sources (G_S1K3) / even-row boxes of the stride-2 programs (same origin as class 0)
                const int x1c = GEO == G_S2K3 ? x0c : (tc.ox0 >> 1) - 4, y1c = GEO == G_S2K3 ? y0c : (tc.oy0 >> 1) - 1;
                for (int s = 0; s < p.n_stages; ++s) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const V2Chunk ch = p.stage[s].ch[h];
                        if (ch.kind == CK_TMA || ch.kind == CK_TMA_P0 || ch.kind == CK_XYZ_T) {
                            const int rslot = (int)raw_r.slot;
                            V2_WAIT_RELAXED(BAR(B_RAW_EMPTY + rslot), raw_r.phase ^ 1u, 64);
                            V2_STAMP(1, dbg_i); ++dbg_i;
                            mbar_arrive_expect_tx(BAR(B_RAW_FULL + rslot), ch.cls ? bytes1 : bytes0);
                            tma_load_4d(smem_u32(sRaw + (size_t)rslot * p.raw_bytes), &maps.m[ch.tmap], ch.cls ? x1c : x0c, ch.cls ? y1c : y0c,
                                        (int)ch.c0, tc.n, BAR(B_RAW_FULL + rslot));
                            raw_r.advance((uint32_t)p.n_raw);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == CW + 1) {
        // =================================================================================================================
        // weight producer (one thread): per (stage, use, tap) one slot = [chunk 0: hi rows, lo rows][chunk 1: hi rows, lo rows]
        // =================================================================================================================
        if (elect_one()) {
            Ring w_r;
            for (long long tile = tile0; tile < p.n_tiles; tile += tstep) {
                const int split = (int)(tile % p.n_split);
                for (int s = 0; s < p.n_stages; ++s) {
                    const V2Stage& st = p.stage[s];
                    for (int u = 0; u < st.n_use;) {
                        const V2Use us = st.use[u];
                        const V2Conv& cv = p.conv[us.conv];
                        const int ntap = __popc((unsigned)us.tapmask);
                        const int Tu = ntap < cv.T ? ntap : cv.T;
                        int n_merge = 0;
                        for (int t0 = 0; t0 < ntap; t0 += Tu, w_r.advance((uint32_t)p.n_w)) {
                            const int nt = ntap - t0 < Tu ? ntap - t0 : Tu;
                            const int ws = (int)w_r.slot;
                            uint32_t total;
                            n_merge = w_slot_merge(p, st, u, t0, nt, ntap, &total);
                            V2_WAIT_RELAXED(BAR(B_W_EMPTY + ws), w_r.phase ^ 1u, 32);
                            uint32_t dst = smem_u32(sW + (size_t)ws * p.w_bytes);
                            mbar_arrive_expect_tx(BAR(B_W_FULL + ws), total);
                            // this use's taps [t0, t0 + nt), then every tap of the merged following uses
                            for (int v = u; v <= u + n_merge; ++v) {
                                const V2Use uv = st.use[v];
                                const V2Conv& cw = p.conv[uv.conv];
                                const uint32_t piece = (uint32_t)cw.Nsl * 64u;              // one tap of this filter-split part
                                const size_t tile_b = (size_t)cw.Npad * 64;
                                const int tv0 = v == u ? t0 : 0, ntv = v == u ? nt : __popc((unsigned)uv.tapmask);
                                const uint8_t* src = cw.wpack + (size_t)(uv.wtile + tv0) * tile_b + (size_t)split * piece;
                                if (p.n_split == 1) {
                                    bulk_g2s(dst, src, (uint32_t)ntv * piece, BAR(B_W_FULL + ws));  // consecutive taps are consecutive tiles
                                } else {
                                    for (int k = 0; k < ntv; ++k) bulk_g2s(dst + (uint32_t)k * piece, src + (size_t)k * tile_b, piece, BAR(B_W_FULL + ws));
                                }
                                dst += (uint32_t)ntv * piece;
                            }
                        }
                        u += 1 + n_merge;
                    }
                }
            }
        }
        __syncwarp();
    } else if (OF && warp == CW + 3) {
        // =================================================================================================================
        // operand-format producer (one thread): chunks whose source already holds bf16 hi / lo 16-byte entries are loaded by TMA
        // STRAIGHT into the operand planes of the A stage (a 5-D box with traversal stride 2 in x and y is one parity plane) --
        // no raw ring, no converter work.  It arrives on A_FULL of every stage (with the byte count of its boxes, or plain).
        // =================================================================================================================
        if (p.has_of && elect_one()) {
            Ring a_r;
            const uint32_t plane_b = (uint32_t)p.plane_px * 16u;
            const uint32_t box = (uint32_t)p.of_box_bytes;
            for (long long tile = tile0; tile < p.n_tiles; tile += tstep) {
                const TileCoord tc = decode_tile(p, tile);
                const int x0c = tc.ox0 - 1, y0c = tc.oy0 - 1;   // G_S2K3: entry (0, 0) of every parity plane = plane pixel (ox0 - 1, oy0 - 1)
                for (int s = 0; s < p.n_stages; ++s) {
                    const V2Stage& st = p.stage[s];
                    const int aslot = (int)a_r.slot;
                    V2_WAIT_RELAXED(BAR(B_A_EMPTY + aslot), a_r.phase ^ 1u, 32);
                    uint32_t nbox = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) nbox += st.ch[h].kind == CK_OF ? 8u : (st.ch[h].kind == CK_OF_P0 ? 2u : 0u);
                    if (nbox) mbar_arrive_expect_tx(BAR(B_A_FULL + aslot), nbox * box);
                    else      mbar_arrive(BAR(B_A_FULL + aslot));
                    const uint32_t A = smem_u32(sA + (size_t)aslot * p.a_bytes);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const V2Chunk ch = st.ch[h];
                        if (ch.kind != CK_OF && ch.kind != CK_OF_P0) continue;
                        // CK_OF: the four parity planes of a full operand-format tensor (plane index = 4 * channel group + parity);
                        // CK_OF_P0: plane 0 -- of a full tensor (cls 0) or of an even-pixel tensor (cls 1: one plane per channel group)
                        const int npl = ch.kind == CK_OF ? 4 : 1;
                        const int pl0 = ch.cls ? (int)(ch.c0 >> 3) : 4 * (int)(ch.c0 >> 3);
                        for (int pl = 0; pl < npl; ++pl) {
                            const uint32_t dst = A + (uint32_t)((pl * 4 + h) * p.plane_px) * 16u;
                            tma_load_5d(dst, &maps.m[ch.tmap], 0, x0c, y0c, pl0 + pl, tc.n, BAR(B_A_FULL + aslot));
                            tma_load_5d(dst + 2u * plane_b, &maps.m[ch.tmap + 1], 0, x0c, y0c, pl0 + pl, tc.n, BAR(B_A_FULL + aslot));
                        }
                    }
                    a_r.advance((uint32_t)p.n_a);
                }
            }
        }
        __syncwarp();
    } else if (warp == CW + 2) {
        // =================================================================================================================
        // MMA issuer.  The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers;
        // one elected lane issues (tcgen05.mma / commit are single-thread instructions).
        // =================================================================================================================
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        Ring a_r, w_r, t_r;
        int dbg_i = 0, dbg_t = 0;
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                        // A: SBO = 128 B, descriptor version 1
        // B: SBO of hi-only rows (512 B) / interleaved [hi|lo] rows of the dual form (256 B)
        const uint32_t b_desc_hi = DUAL ? ((256u >> 4) | (1u << 14)) : ((512u >> 4) | (1u << 14));
        const uint32_t b_lbo = (128u >> 4) << 16;                                  // B: LBO = 128 B between the two K chunks
        const uint32_t a_lbo = (((uint32_t)p.plane_px * 16u) >> 4) << 16;          // LBO: chunk 0 -> chunk 1 plane
        const uint32_t a_lo_off = (2u * (uint32_t)p.plane_px * 16u) >> 4;          // hi planes -> lo planes
        const uint32_t tile_cols = (uint32_t)(p.MB * p.cols_block);
        const uint32_t sA16 = smem_u32(sA) >> 4, sW16 = smem_u32(sW) >> 4;
        const uint32_t a_bytes16 = (uint32_t)p.a_bytes >> 4, w_bytes16 = (uint32_t)p.w_bytes >> 4;
        const uint32_t n_w = (uint32_t)p.n_w, n_a = (uint32_t)p.n_a, acc_bufs = (uint32_t)p.acc_bufs;
        const int MB = p.MB, n_cmds = p.n_cmds;
        const uint32_t cols_block = (uint32_t)p.cols_block;
        // ONE elected lane runs the whole role: per-block elect/reconvergence costs ~190 clk on this path (tools/umma_rate_probe.cu).
        // The per-tap work is a table walk: the stage / use / tap / weight-slot structure of the program was flattened on the host
        // (build_commands), because this single thread's scalar bookkeeping -- not the tensor pipe -- set the pace of the low-N layers
        // (~60 instructions per tap for three MMAs; profiles/r02b_kb_timeline.txt).
        if (elect_one()) {
            for (long long tile = tile0; tile < p.n_tiles; tile += tstep) {
                mbar_wait(BAR(B_ACC_EMPTY + t_r.slot), t_r.phase ^ 1u);
                tc_fence_after();
                V2_STAMP(4, dbg_t); ++dbg_t;
                const uint32_t acc = tmem_u + t_r.slot * tile_cols;
                uint32_t a_base16 = 0, b_base16 = 0;
                uint4 nxt = p.cmd[0];
#pragma unroll 1
                for (int i = 0; i < n_cmds; ++i) {
                    const uint4 c = nxt;
                    nxt = p.cmd[i + 1];
                    const uint32_t flags = c.x >> 16;
                    if (flags & CMD_WAIT_A) {
                        mbar_wait(BAR(B_A_FULL + a_r.slot), a_r.phase);
                        tc_fence_after();
                        V2_STAMP(2, dbg_i); ++dbg_i;
                        a_base16 = (sA16 + a_r.slot * a_bytes16) | a_lbo;
                    }
                    if (flags & CMD_WAIT_W) {
                        mbar_wait(BAR(B_W_FULL + w_r.slot), w_r.phase);  // weights land through the async proxy: the barrier's acquire is enough
                        b_base16 = (sW16 + w_r.slot * w_bytes16) | b_lbo;
                    }
                    uint32_t a0 = a_base16 + (c.x & 0xFFFFu);
                    const uint32_t b16 = b_base16 + (c.y & 0xFFFFu);
                    uint32_t d = acc + (c.y >> 16);
                    const uint32_t idesc = c.z;
                    const uint32_t accum = flags & CMD_ACCUM;
#ifndef KBNET_V2_KO_MMA   // (knock-out experiments: which role bounds a tile?  results are garbage with these flags)
                    if (DUAL) {  // B' = [b_hi | b_lo] interleaved per 8-filter group: all four hi/lo products in two MMAs
#pragma unroll 1
                        for (int mb = 0; mb < MB; ++mb, a0 += 128u, d += cols_block) {
                            mma_bf16_ss(d, make_desc(a0, desc_hi), make_desc(b16, b_desc_hi), idesc, accum);
                            mma_bf16_ss(d, make_desc(a0 + a_lo_off, desc_hi), make_desc(b16, b_desc_hi), idesc, 1u);
                        }
                    } else {
#pragma unroll 1
                        for (int mb = 0; mb < MB; ++mb, a0 += 128u, d += cols_block) {
                            mma_bf16_ss(d, make_desc(a0, desc_hi), make_desc(b16, b_desc_hi), idesc, accum);
                            mma_bf16_ss(d, make_desc(a0 + a_lo_off, desc_hi), make_desc(b16, b_desc_hi), idesc, 1u);
                            mma_bf16_ss(d, make_desc(a0, desc_hi), make_desc(b16 + 16u, b_desc_hi), idesc, 1u);
                        }
                    }
#else
                    (void)a0; (void)b16; (void)d; (void)idesc; (void)accum;
#endif
                    if (flags & CMD_COMMIT_W) {
                        mma_commit(BAR(B_W_EMPTY + w_r.slot));
                        w_r.advance(n_w);
                    }
                    if (flags & CMD_COMMIT_A) {
                        mma_commit(BAR(B_A_EMPTY + a_r.slot));
                        V2_STAMP(2, dbg_i); ++dbg_i;
                        a_r.advance(n_a);
                    }
                }
                mma_commit(BAR(B_ACC_FULL + t_r.slot));
                t_r.advance(acc_bufs);
            }
        }
        __syncwarp();
    } else if (warp >= CW + 4) {
        // =================================================================================================================
        // epilogue: TMEM -> registers -> activation -> NCHW
        // =================================================================================================================
        const int quarter = warp & 3;
        const uint32_t tile_cols = (uint32_t)(p.MB * p.cols_block);
        Ring t_r;
        int dbg_i = 0;
        for (long long tile = tile0; tile < p.n_tiles; tile += tstep) {
            const TileCoord tc = decode_tile(p, tile);
            const uint32_t buf = t_r.slot;
            V2_WAIT_EPI(BAR(B_ACC_FULL + buf), t_r.phase);
            if (tid == 12 * 32) { V2_STAMP(3, dbg_i); ++dbg_i; }
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * tile_cols + ((uint32_t)(quarter * 32) << 16);
            for (int mb = 0; mb < p.MB; ++mb) {
                const int q = mb * 128 + quarter * 32 + lane;
                const int oyl = q / p.P, oxl = q - oyl * p.P;
                const int oy = tc.oy0 + oyl, ox = tc.ox0 + oxl;
                const bool valid = oyl < p.R && oxl < p.TW && oy < p.Hout && ox < p.Wout;
                for (int j = 0; j < p.n_convs; ++j) {
                    const V2Conv& cv = p.conv[j];
                    const int f0 = tc.split * cv.Nsl;
                    const int nf = cv.F - f0 < cv.Nsl ? cv.F - f0 : cv.Nsl;  // real filters of this split part (may be <= 0)
                    if (nf <= 0) continue;
                    // warp-uniform 64-bit base + per-thread 32-bit element offset (one sample's output is far below 2^32 elements):
                    // one IMAD.WIDE per store instead of a 64-bit multiply-add chain
                    // running pointer: one 64-bit add per store instead of a multiply-add chain per filter plane
                    const size_t oplane = (size_t)cv.oplane;
                    float* dst = cv.out + (size_t)tc.n * cv.out_stride_n + (size_t)f0 * oplane + (size_t)cv.o_off + (size_t)oy * cv.o_row_pitch +
                                 (size_t)ox * cv.o_col_mul;
                    const uint32_t tcol = acc + (uint32_t)(mb * p.cols_block + cv.col);
                    const bool act = cv.act != 0;
                    for (int c = 0; c < nf; c += 16) {
                        float o[16];
                        if (DUAL) {  // filter 8g + j lives in columns 16g + j (b_hi products) and 16g + 8 + j (b_lo products)
                            uint32_t r0[16], r1[16];
                            tmem_ld_x16(tcol + (uint32_t)(2 * c), r0);
                            tmem_ld_x16(tcol + (uint32_t)(2 * c + 16), r1);
                            tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                o[k] = __uint_as_float(r0[k]) + __uint_as_float(r0[k + 8]);
                                o[k + 8] = __uint_as_float(r1[k]) + __uint_as_float(r1[k + 8]);
                            }
                        } else {
                            uint32_t r0[16];
                            tmem_ld_x16(tcol + (uint32_t)c, r0);
                            tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 16; ++k) o[k] = __uint_as_float(r0[k]);
                        }
                        if (act) {
#pragma unroll
                            for (int k = 0; k < 16; ++k) o[k] = leaky(o[k], p.slope);
                        }
#ifdef KBNET_V2_KO_EPI
                        if (valid && cv.out != nullptr && o[0] == 12345.678f) {
#else
                        if (valid && (!OF || cv.out != nullptr)) {
#endif
                            float* d = dst;
                            if (c + 16 <= nf) {
#pragma unroll
                                for (int k = 0; k < 16; ++k, d += oplane) *d = o[k];
                            } else {
#pragma unroll
                                for (int k = 0; k < 16; ++k, d += oplane)
                                    if (c + k < nf) *d = o[k];
                            }
                        }
                        // operand-format copy: 8 filters of this pixel = one 16-byte entry (hi) + one (lo) in the pixel's parity plane
                        if (OF && valid && cv.of_hi != nullptr && !(cv.of_even && ((oy | ox) & 1))) {
                            const size_t of_plane = (size_t)(p.Hout >> 1) * (size_t)(p.Wout >> 1);
                            const size_t pix = (size_t)(oy >> 1) * (size_t)(p.Wout >> 1) + (size_t)(ox >> 1);
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                if (c + 8 * hf < nf) {
                                    const size_t cgi = (size_t)tc.n * cv.of_cg + (size_t)((f0 + c) >> 3) + hf;
                                    const size_t e = (cv.of_even ? cgi * of_plane + pix : (cgi * 4 + (size_t)((oy & 1) * 2 + (ox & 1))) * of_plane + pix) * 16;
                                    store_split8(cv.of_hi + e, cv.of_lo + e, reinterpret_cast<const float(&)[8]>(o[8 * hf]));
                                }
                            }
                        }
                        dst += 16 * oplane;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + buf));
            if (tid == 12 * 32) { V2_STAMP(3, dbg_i); ++dbg_i; }
            t_r.advance((uint32_t)p.acc_bufs);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CW + 2) tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
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

inline int roundup(int a, int b) { return (a + b - 1) / b * b; }

// an NCHW fp32 source a 4-D tiled tensor map can describe
bool tma_ok(const ConvSrcDev& s) {
    return s.ptr != nullptr && ((uintptr_t)s.ptr & 15) == 0 && (s.W & 3) == 0 && (s.stride_n & 3) == 0 && s.W >= 4;
}

// row_step = 2: the box visits every other row (TMA traversal stride): bh rows are LOADED, spanning 2 * bh source rows
bool make_map4(CUtensorMap* tm, const ConvSrcDev& s, int N, int bw, int bh, int row_step = 1) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || bw > 256 || bh * row_step > 256 || bw < 1 || bh < 1) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.C, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)s.W * 4, (cuuint64_t)s.W * s.H * 4, (cuuint64_t)s.stride_n * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)(bh * row_step), 8, 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)row_step, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(s.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// operand-format source (operand_format.cu): `planes` = 4 * channel groups (full tensor) or the channel groups (even-pixel tensor)
// parity planes of H2 x W2 16-byte entries per sample: 5-D map {8 bf16, W2, H2, planes, N}; one box = P x R1 entries of one plane
bool make_map_of(CUtensorMap* tm, const void* ptr, int N, int planes, int H2, int W2, int P, int R1) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || ptr == nullptr || ((uintptr_t)ptr & 15) || P > 256 || R1 > 256 || planes < 1 || H2 < 1 || W2 < 1) return false;
    const cuuint64_t dims[5] = {8, (cuuint64_t)W2, (cuuint64_t)H2, (cuuint64_t)planes, (cuuint64_t)N};
    const cuuint64_t strides[4] = {16, (cuuint64_t)W2 * 16, (cuuint64_t)W2 * H2 * 16, (cuuint64_t)W2 * H2 * planes * 16};
    const cuuint32_t box[5] = {8, (cuuint32_t)P, (cuuint32_t)R1, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

double mma_cycles(int N) {  // measured with a lean single-lane issue loop, tools/umma_rate_probe.cu (profiles/r01_umma_rate_probe.log)
    const double c = N / 2.0, s = 44.0 + N / 6.0;
    return c > s ? c : s;
}

struct V2Plan {  // what the tile search decides
    int R, TW, n_split, dual;
    int wcap;  // bytes of weights per ring slot (0 = the default 24 KB); only the measured selection varies it
    int ring;  // which ring gets the spare shared memory first: 0 default order, 1 operand stages, 2 weights, 3 raw tiles (measured selection)
};

// Fill the geometry-dependent fields of p for a candidate (R, TW, n_split); returns the shared-memory footprint or -1.
long long apply_geometry(V2Params& p, int R, int TW, int n_split, int dual, bool has_up, int plan_wcap = 0) {
    const int geo = p.geo;
    p.R = R; p.TW = TW; p.n_split = n_split; p.dual = dual;
    int max_shift;
    if (geo == G_S1K3) { p.P = TW + 2; p.n_entries = (R + 2) * p.P; max_shift = 2 * p.P + 2; p.BW[0] = roundup(p.P + 3, 4); p.BH[0] = R + 2; }
    else if (geo == G_S1K1) { p.P = TW; p.n_entries = R * p.P; max_shift = 0; p.BW[0] = TW; p.BH[0] = R; }
    else if (geo == G_S2K3) { p.P = TW + 1; p.n_entries = (R + 1) * p.P; max_shift = p.P + 1; p.BW[0] = roundup(2 * p.P + 2, 4); p.BH[0] = 2 * R + 2; }
    else { p.P = TW; p.n_entries = R * p.P; max_shift = 0; p.BW[0] = 2 * TW; p.BH[0] = 2 * R; }
    if (p.has_of && (geo != G_S2K3 || p.P > 256 || R + 1 > 256)) return -1;   // one operand-format box = one parity plane
    p.of_box_bytes = p.P * (R + 1) * 16;
    p.BW[1] = p.BH[1] = 4;
    if (has_up) { p.BW[1] = roundup(TW / 2 + 5, 4); p.BH[1] = R / 2 + 2; }
    if (geo == G_S2K3) { p.BW[1] = p.BW[0]; p.BH[1] = R + 1; }  // class 1 of the stride-2 programs: even rows only (sources read at even pixels)
    p.n_planes = geo == G_S2K3 ? 4 : 1;
    p.MB = cdiv(R * p.P, 128);
    const int max_entries = (geo == G_S2K3 ? 1 : (geo == G_S2K1 ? 2 : 4)) * conv_warps(geo) * 32;
    if (p.n_entries > max_entries || p.BW[0] > 256 || p.BH[0] > 256 || p.BW[1] > 256 || p.BH[1] > 256) return -1;
    p.plane_px = roundup(p.n_entries > max_shift + p.MB * 128 ? p.n_entries : max_shift + p.MB * 128, 8);
    int cols = 0, max_nsl = 0;
    for (int j = 0; j < p.n_convs; ++j) {
        V2Conv& cv = p.conv[j];
        if (cv.Npad % (16 * n_split)) return -1;
        cv.Nsl = cv.Npad / n_split;
        if (cv.Nsl > 256 || (dual && cv.Nsl > 128)) return -1;
        cv.col = cols;
        cols += (dual ? 2 : 1) * cv.Nsl;
        max_nsl = cv.Nsl > max_nsl ? cv.Nsl : max_nsl;
    }
    p.cols_block = cols;
    if (p.MB * cols > 512) return -1;
    p.acc_bufs = 2 * p.MB * cols <= 512 ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.acc_bufs * p.MB * cols) p.tmem_cols <<= 1;
    const int raw0 = p.BW[0] * p.BH[0] * 32, raw1 = has_up ? p.BW[1] * p.BH[1] * 32 : 0;
    p.raw_bytes = roundup(raw0 > raw1 ? raw0 : raw1, 128);
    p.a_bytes = roundup(p.n_planes * 4 * p.plane_px * 16, 128);
    // weight-ring slot: as many taps of a 3x3 use as fit ~24 KB (one barrier round trip and one bulk copy per slot)
    static const int env_wcap = getenv("KBNET_V2_WCAP") ? atoi(getenv("KBNET_V2_WCAP")) : 0;  // debugging aid: overrides the plan
    const int wcap = env_wcap > 0 ? env_wcap : (plan_wcap > 0 ? plan_wcap : 24 * 1024);      // bytes of weights per ring slot
    p.T = max_nsl * 64 * 9 <= wcap ? 9 : (max_nsl * 64 * 4 <= wcap ? 4 : (max_nsl * 64 * 3 <= wcap ? 3 : (max_nsl * 64 * 2 <= wcap ? 2 : 1)));
    int max_taps = 1;
    for (int j = 0; j < p.n_convs; ++j) max_taps = p.conv[j].ntaps > max_taps ? p.conv[j].ntaps : max_taps;
    if (p.T > max_taps) p.T = max_taps;
    p.w_bytes = roundup(max_nsl * 64 * p.T, 128);
    if (p.n_convs > 1) {   // several small uses per stage (the four parity convolutions of a 64 -> 12 up-convolution): room for a whole stage
        int stage_max = 0;
        for (int s = 0; s < p.n_stages; ++s) {
            int sum = 0;
            for (int u = 0; u < p.stage[s].n_use; ++u)
                sum += __builtin_popcount((unsigned)p.stage[s].use[u].tapmask) * p.conv[p.stage[s].use[u].conv].Nsl * 64;
            stage_max = sum > stage_max ? sum : stage_max;
        }
        if (stage_max <= wcap && roundup(stage_max, 128) > p.w_bytes) p.w_bytes = roundup(stage_max, 128);
    }
    for (int j = 0; j < p.n_convs; ++j) {   // per-convolution taps per slot: fewer slot hand-overs for the convolutions with few filters
        V2Conv& cv = p.conv[j];
        const int fit = p.w_bytes / (cv.Nsl * 64);
        cv.T = fit >= 9 ? 9 : (fit >= 4 ? 4 : (fit >= 3 ? 3 : (fit >= 2 ? 2 : 1)));
        if (cv.T < p.T) cv.T = p.T;
    }
    p.tiles_x = cdiv(p.Wout, TW);
    p.tiles_y = cdiv(p.Hout, R);
    p.n_tiles = (long long)p.N * p.tiles_x * p.tiles_y * n_split;
    // minimum rings: 2 raw, 2 A, 2 W
    return (long long)V2_HEADER + 2LL * p.raw_bytes + 2LL * p.a_bytes + 2LL * p.w_bytes;
}

// Memo of the tile search (a pure function of the program shape); guarded by a mutex, so concurrent callers are fine.
struct PlanKey {
    int v[24];
    bool operator==(const PlanKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct PlanEntry {
    PlanKey key;
    V2Plan plan;
};
std::mutex g_plan_mutex;
std::vector<PlanEntry> g_plans;

// Candidate plans ordered by the cost model: the best (R) for every (filter split, split mode, tile width), cheapest first.
struct Candidate {
    V2Plan plan;
    double cost;
};
std::vector<Candidate> search_plans(const V2Params& p, bool has_up, int n_chunks, int n_tma_chunks, int sms) {
    const int geo = p.geo;
    std::vector<Candidate> cands;
    static thread_local V2Params trial;
    trial = p;
    const int tw_step = geo == G_S1K3 ? (has_up ? 8 : 4) : (geo == G_S1K1 ? 4 : 2);
    const int r_step = (geo == G_S1K3 && has_up) ? 2 : 1;
    for (int n_split = 1; n_split <= 8; n_split *= 2) {
        for (int dual = 1; dual >= 0; --dual) {
            // per-M-block MMA cycles and per-tile weight bytes of this (n_split, dual)
            double mma_mb = 0, wbytes = 0;
            bool ok = true;
            for (int j = 0; j < p.n_convs; ++j)
                if (p.conv[j].Npad % (16 * n_split) || p.conv[j].Npad / n_split > (dual ? 128 : 256)) ok = false;
            if (!ok) continue;
            for (int s = 0; s < p.n_stages; ++s)
                for (int u = 0; u < p.stage[s].n_use; ++u) {
                    const V2Use& us = p.stage[s].use[u];
                    const int nsl = p.conv[us.conv].Npad / n_split;
                    const int nt = __builtin_popcount((unsigned)us.tapmask);
                    mma_mb += nt * (dual ? 2.0 * mma_cycles(2 * nsl) : 3.0 * mma_cycles(nsl));
                    wbytes += (double)nt * nsl * 64.0;
                }
            int last_tw = -1;
            for (int parts = 1; parts <= 64; ++parts) {
                int TW = roundup(cdiv(p.Wout, parts), tw_step);
                if (TW > 248) continue;
                if (TW == last_tw) continue;
                last_tw = TW;
                double best = 1e300;
                V2Plan plan = {0, 0, 0, 0, 0, 0};
                for (int R = r_step; R <= 64; R += r_step) {
                    if (R - r_step >= p.Hout) break;
                    const long long smem_min = apply_geometry(trial, R, TW, n_split, dual, has_up);
                    if (smem_min < 0 || smem_min > V2_SMEM_LIMIT) continue;
                    const double mma = mma_mb * trial.MB;
                    const double items = (double)cdiv(trial.n_entries, conv_warps(geo) * 32);
                    const double cvt = (double)n_chunks * items * (trial.n_planes * 45.0 + 10.0) * conv_warps(geo) / 2.0 + p.n_stages * 150.0;
                    double outb = 0;
                    for (int j = 0; j < p.n_convs; ++j) outb += (double)trial.conv[j].Nsl * R * TW * 4.0;
                    const double mem = ((double)n_tma_chunks * trial.raw_bytes + wbytes + outb) / 24.0;  // ~ one SM's share of HBM/L2, B/clk
                    const double epi = (double)trial.MB * trial.cols_block * 10.0;                       // tcgen05.ld + stores of 4 warps
                    double tile = mma > cvt ? mma : cvt;
                    tile = tile > mem ? tile : mem;
                    tile = tile > epi ? tile : epi;
                    if (trial.acc_bufs == 1) tile += epi;
                    const double waves = (double)cdiv((int)(trial.n_tiles > (1 << 30) ? (1 << 30) : trial.n_tiles), sms);
                    const double total = waves * tile + tile + 4000.0;  // + pipeline fill
                    if (total < best) { best = total; plan = {R, TW, n_split, dual, 0, 0}; }
                }
                if (plan.R > 0) cands.push_back({plan, best});
            }
        }
    }
    std::sort(cands.begin(), cands.end(), [](const Candidate& a, const Candidate& b) { return a.cost < b.cost; });
    return cands;
}

std::atomic<int> g_autotune{-1};  // -1: not initialised (KBNET_AUTOTUNE), 0 off, 1 on

// Flatten the program (stages x uses x taps, weight-ring slots of p.T taps) into the MMA issuer's command table for the geometry
// apply_geometry() has just set.  The weight producer walks the same (stage, use, slot) order.
int build_commands(V2Params& p) {
    const int geo = p.geo;
    const uint32_t P_u = (uint32_t)p.P, plane4 = 4u * (uint32_t)p.plane_px;
    // operand-plane offset (in 16-byte rows) of tap (ky, kx): a 3x3 tap is the same plane read from a shifted start address
    auto tap_offset = [&](uint32_t ky, uint32_t kx) -> uint32_t {
        if (geo == G_S1K3) return ky * P_u + kx;  // (tap 4 = the centre: P + 1)
        if (geo == G_S2K3)  // dy = ky - 1, dx = kx - 1: parity plane (dy != 0, dx != 0), shift (dy >= 0) * P + (dx >= 0)
            return ((ky != 1u ? 2u : 0u) + (kx != 1u ? 1u : 0u)) * plane4 + (ky != 0u ? P_u : 0u) + (kx != 0u ? 1u : 0u);
        return 0u;
    };
    int n = 0;
    for (int s = 0; s < p.n_stages; ++s) {
        const V2Stage& st = p.stage[s];
        const int first_cmd = n;
        int merged_left = 0;          // uses still to come that share the currently open weight slot
        uint32_t slot_off16 = 0;      // where the next merged use starts inside that slot (16-byte units)
        for (int u = 0; u < st.n_use; ++u) {
            const V2Use& us = st.use[u];
            const V2Conv& cv = p.conv[us.conv];
            const uint32_t piece16 = (uint32_t)cv.Nsl * 4u;   // one tap of weights, in 16-byte units
            const uint32_t idesc = idesc_bf16_f32(128, p.dual ? 2 * cv.Nsl : cv.Nsl);
            const int ntap = __builtin_popcount((unsigned)us.tapmask);
            int left = ntap;
            const int Tu = left < cv.T ? left : cv.T;
            const bool follower = merged_left > 0;   // this use's (single) slot is the tail of the previous use's last slot
            if (follower) --merged_left;
            int cnt = 0, t0 = 0;
            uint32_t accum = us.first ? 0u : CMD_ACCUM;
            for (uint32_t t = 0; t < 9; ++t) {
                if (!((us.tapmask >> t) & 1u)) continue;
                KB_CHECK_ARG(n < V2_MAX_CMDS, "conv(v2): program too long for the command table");
                uint32_t flags = accum | ((cnt == 0 && !follower) ? CMD_WAIT_W : 0u);
                if (cnt == 0 && !follower) {   // a slot opens: does it take following uses?
                    const int nt = ntap - t0 < Tu ? ntap - t0 : Tu;
                    uint32_t total;
                    merged_left = w_slot_merge(p, st, u, t0, nt, ntap, &total);
                    slot_off16 = (uint32_t)nt * piece16;
                }
                const uint32_t a_rel = tap_offset(t / 3, t % 3), b_rel = (follower ? slot_off16 : 0u) + (uint32_t)cnt * piece16;
                KB_CHECK_ARG(a_rel < 0x10000u && b_rel < 0x10000u && cv.col < 0x10000, "conv(v2): command field overflow");
                accum = CMD_ACCUM;
                --left;
                if (++cnt == Tu || left == 0) {
                    // the slot closes with its last occupant: this use, unless followers share it
                    if (!(left == 0 && merged_left > 0)) flags |= CMD_COMMIT_W;
                    t0 += cnt;
                    cnt = 0;
                }
                p.cmd[n++] = make_uint4(a_rel | (flags << 16), b_rel | ((uint32_t)cv.col << 16), idesc, 0u);
            }
            if (follower) slot_off16 += (uint32_t)ntap * piece16;
        }
        KB_CHECK_ARG(n > first_cmd, "conv(v2): stage %d has no taps", s);
        p.cmd[first_cmd].x |= (uint32_t)CMD_WAIT_A << 16;
        p.cmd[n - 1].x |= (uint32_t)CMD_COMMIT_A << 16;
    }
    p.cmd[n] = make_uint4(0u, 0u, 0u, 0u);
    p.n_cmds = n;
    return 0;
}

// geometry + rings + tensor maps + launch for one plan
struct OfSrc {   // operand-format tensor: hi and lo halves of the bf16 split, `planes` parity planes of H2 x W2 entries per sample
    const void* hi;
    const void* lo;
    int planes, H2, W2;
};

int launch_with_plan(V2Params& p, const V2Plan& plan, const ConvSrcDev* map_src, int n_maps, const uint8_t* map_cls, bool has_up, int sms,
                     cudaStream_t stream, bool report, const OfSrc* of_src = nullptr, int n_of = 0) {
    const int geo = p.geo;
    long long used = apply_geometry(p, plan.R, plan.TW, plan.n_split, plan.dual, has_up, plan.wcap);
    KB_CHECK_ARG(used > 0 && used <= V2_SMEM_LIMIT, "conv(v2): plan R=%d TW=%d split=%d dual=%d does not fit", plan.R, plan.TW, plan.n_split, plan.dual);
    if (int rc = build_commands(p)) return rc;
    p.n_raw = 2; p.n_a = 2; p.n_w = 2;
    // Spend the remaining shared memory on ring depth, in the order in which exposed latency was measured to cost (role timelines,
    // tools/v2_timeline.py): a raw slot is refilled only after the converters release it, so with 2 slots every chunk waits a full
    // TMA round trip (~2 k clk); weights and a third operand stage come next.
    auto grow = [&](int& n, int bytes, int upto) {
        while (n < upto && used + bytes <= V2_SMEM_LIMIT) { ++n; used += bytes; }
    };
    if (plan.ring == 1 && p.n_stages > 2) grow(p.n_a, p.a_bytes, 3);
    if (plan.ring == 2) grow(p.n_w, p.w_bytes, 4);
    if (plan.ring == 3) grow(p.n_raw, p.raw_bytes, 4);
    grow(p.n_raw, p.raw_bytes, 3);
    grow(p.n_w, p.w_bytes, 3);
    grow(p.n_raw, p.raw_bytes, 4);
    if (p.n_stages > 2) grow(p.n_a, p.a_bytes, 3);
    grow(p.n_w, p.w_bytes, 4);
    grow(p.n_raw, p.raw_bytes, V2_MAX_RAW);
    grow(p.n_w, p.w_bytes, V2_MAX_W);
    if (const char* ov = getenv("KBNET_V2_RINGS")) {  // debugging aid: "n_raw,n_a,n_w" (must fit shared memory)
        int a, b, c;
        if (sscanf(ov, "%d,%d,%d", &a, &b, &c) == 3 && a >= 2 && a <= V2_MAX_RAW && b >= 2 && b <= V2_MAX_A && c >= 2 && c <= V2_MAX_W) {
            const long long u2 = (long long)V2_HEADER + (long long)a * p.raw_bytes + (long long)b * p.a_bytes + (long long)c * p.w_bytes;
            if (u2 <= V2_SMEM_LIMIT) { p.n_raw = a; p.n_a = b; p.n_w = c; used = u2; }
        }
    }
    if (used < 120 * 1024) used = 120 * 1024;  // one CTA per SM (the TMEM budget assumes it)

    V2Maps maps;
    memset(&maps, 0, sizeof(maps));
    for (int i = 0; i < n_maps && i < 4; ++i)
        KB_CHECK_ARG(make_map4(&maps.m[i], map_src[i], p.N, p.BW[map_cls[i]], p.BH[map_cls[i]], (geo == G_S2K3 && map_cls[i] == 1) ? 2 : 1),
                     "conv(v2): cuTensorMapEncodeTiled failed for source %d", i);

    for (int i = 0; i < n_of; ++i) {
        KB_CHECK_ARG(4 + 2 * i + 1 < V2_MAX_MAPS, "conv(v2): too many operand-format sources");
        KB_CHECK_ARG(make_map_of(&maps.m[4 + 2 * i], of_src[i].hi, p.N, of_src[i].planes, of_src[i].H2, of_src[i].W2, p.P, p.R + 1) &&
                     make_map_of(&maps.m[5 + 2 * i], of_src[i].lo, p.N, of_src[i].planes, of_src[i].H2, of_src[i].W2, p.P, p.R + 1),
                     "conv(v2): cuTensorMapEncodeTiled failed for operand-format source %d", i);
    }

    const long long grid = p.n_tiles < sms ? p.n_tiles : sms;
    KB_CHECK_ARG(grid > 0, "conv(v2): empty grid");
    static const char* timeline = getenv("KBNET_V2_TIMELINE");  // debugging aid: dump clock64 stamps of CTA 0 to this file (synchronises!)
    p.dbg = nullptr;
    if (timeline && report) {
        KB_CUDA(cudaMalloc(&p.dbg, 6 * 4096 * sizeof(long long)));
        KB_CUDA(cudaMemsetAsync(p.dbg, 0, 6 * 4096 * sizeof(long long), stream));
    }
    static const bool debug = getenv("KBNET_V2_DEBUG") != nullptr;
    if (debug && report)
        fprintf(stderr, "[conv v2] geo %d  %dx%dx%d -> %dx%d  stages %d convs %d Npad %d/%d/%d | R %d TW %d P %d MB %d split %d dual %d T %d | tiles %lld grid %lld | "
                "raw %d x %d B, A %d x %d B, W %d x %d B, smem %lld | tmem %d cols x %d bufs\n",
                geo, p.N, p.Hin, p.Win, p.Hout, p.Wout, p.n_stages, p.n_convs, p.conv[0].Npad, p.conv[1].Npad, p.conv[2].Npad + p.conv[3].Npad, p.R, p.TW, p.P, p.MB,
                p.n_split, p.dual, p.T, p.n_tiles, grid, p.n_raw, p.raw_bytes, p.n_a, p.a_bytes, p.n_w, p.w_bytes, used, p.MB * p.cols_block, p.acc_bufs);
#define V2_LAUNCH_K(KERN, G)                                                                                   \
    do {                                                                                                       \
        KB_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)used));           \
        KERN<<<(unsigned)grid, cta_threads(G), (size_t)used, stream>>>(p, maps);                               \
    } while (0)
#define V2_LAUNCH(G)                                                      \
    do {                                                                  \
        if (p.dual) V2_LAUNCH_K((conv_v2_kernel<G, true, false>), G);     \
        else        V2_LAUNCH_K((conv_v2_kernel<G, false, false>), G);    \
    } while (0)
    bool any_of_out = false;   // an operand-format copy of some output, or operand-format sources: the OF instantiation
    for (int j = 0; j < p.n_convs; ++j) any_of_out = any_of_out || p.conv[j].of_hi != nullptr || p.conv[j].out == nullptr;
    if (geo == G_S2K3 && (p.has_of || any_of_out)) {
        if (p.dual) V2_LAUNCH_K((conv_v2_kernel<G_S2K3, true, true>), G_S2K3);
        else        V2_LAUNCH_K((conv_v2_kernel<G_S2K3, false, true>), G_S2K3);
    } else
    if (geo == G_S1K3) V2_LAUNCH(G_S1K3);
    else if (geo == G_S1K1) V2_LAUNCH(G_S1K1);
    else if (geo == G_S2K3) V2_LAUNCH(G_S2K3);
    else V2_LAUNCH(G_S2K1);
#undef V2_LAUNCH_K
#undef V2_LAUNCH
    KB_CUDA(cudaGetLastError());
    count_launch();
    count_launch_v2();
    if (p.dbg) {
        std::vector<long long> h(6 * 4096);
        KB_CUDA(cudaStreamSynchronize(stream));
        KB_CUDA(cudaMemcpy(h.data(), p.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.dbg);
        p.dbg = nullptr;
        if (FILE* f = fopen(timeline, "w")) {
            fprintf(f, "# stages %d MB %d n_a %d n_raw %d n_w %d acc_bufs %d tiles %lld grid %lld\n", p.n_stages, p.MB, p.n_a, p.n_raw, p.n_w, p.acc_bufs,
                    p.n_tiles, grid);
            for (int r = 0; r < 6; ++r)
                for (int i = 0; i < 4096 && h[r * 4096 + i]; ++i) fprintf(f, "%d %d %lld\n", r, i, h[r * 4096 + i]);
            fclose(f);
        }
    }
    return 0;
}

int v2_launch(V2Params& p, const ConvSrcDev* map_src, int n_maps, const uint8_t* map_cls, bool has_up, cudaStream_t stream,
              const OfSrc* of_src = nullptr, int n_of = 0) {
    const int geo = p.geo;
    int n_ops = 0, n_tma_chunks = 0, n_chunks = 0, n_of_chunks = 0;  // per tile
    for (int s = 0; s < p.n_stages; ++s) {
        for (int h = 0; h < 2; ++h) {
            const int k = p.stage[s].ch[h].kind;
            if (k == CK_OF || k == CK_OF_P0) { ++n_of_chunks; continue; }   // no converter work, no raw slot
            n_chunks += 1;
            if (k == CK_TMA || k == CK_TMA_P0 || k == CK_XYZ_T) ++n_tma_chunks;
        }
        for (int u = 0; u < p.stage[s].n_use; ++u) n_ops += __builtin_popcount((unsigned)p.stage[s].use[u].tapmask);
    }
    p.has_of = n_of_chunks > 0;
    KB_CHECK_ARG(!p.has_of || (geo == G_S2K3 && of_src != nullptr && n_of > 0), "conv(v2): operand-format chunks need a stride-2 3x3 program and their sources");
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            int v = 0;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        }
    }
    if (g_autotune.load() < 0) g_autotune.store(getenv("KBNET_AUTOTUNE") ? (atoi(getenv("KBNET_AUTOTUNE")) != 0) : 0);
    const int autotune = g_autotune.load();
    PlanKey key;
    memset(&key, 0, sizeof(key));
    {
        int* k = key.v;
        k[0] = geo; k[1] = p.N; k[2] = p.Hin; k[3] = p.Win; k[4] = p.Hout; k[5] = p.Wout; k[6] = p.n_stages; k[7] = p.n_convs;
        for (int j = 0; j < p.n_convs; ++j) { k[8 + j] = p.conv[j].Npad; k[12 + j] = p.conv[j].ntaps; }
        k[16] = has_up; k[17] = n_chunks; k[18] = n_tma_chunks; k[19] = n_ops; k[20] = sms; k[21] = autotune != 0; k[22] = n_of_chunks;
    }
    V2Plan plan = {0, 0, 0, 0, 0, 0};
    bool found = false;
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        for (const PlanEntry& e : g_plans)
            if (e.key == key) { plan = e.plan; found = true; break; }
    }
    if (!found) {
        const std::vector<Candidate> cands = search_plans(p, has_up, n_chunks, n_tma_chunks, sms);
        KB_CHECK_ARG(!cands.empty(), "conv(v2): no tile fits");
        plan = cands[0].plan;
        if (autotune && cands.size() > 1) {
            // Measured tile selection (opt-in: it synchronises the stream the first time a shape is seen): run the cost model's
            // best candidates on the real tensors -- every candidate writes the same result -- and keep the fastest.
            cudaEvent_t e0, e1;
            KB_CUDA(cudaEventCreate(&e0));
            KB_CUDA(cudaEventCreate(&e1));
            float best_ms = 1e30f;
            const size_t n_try = cands.size() < 12 ? cands.size() : 12;
            for (size_t i = 0; i < n_try; ++i) {
                int rc = launch_with_plan(p, cands[i].plan, map_src, n_maps, map_cls, has_up, sms, stream, false, of_src, n_of);  // warm-up (I-cache, L2)
                if (rc) continue;
                float cand_ms = 1e30f;
                for (int rep = 0; rep < 3 && !rc; ++rep) {  // best of three: a single shot is too noisy to rank close candidates
                    KB_CUDA(cudaEventRecord(e0, stream));
                    rc = launch_with_plan(p, cands[i].plan, map_src, n_maps, map_cls, has_up, sms, stream, false, of_src, n_of);
                    KB_CUDA(cudaEventRecord(e1, stream));
                    if (rc) break;
                    KB_CUDA(cudaEventSynchronize(e1));
                    float ms = 0.f;
                    KB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                    cand_ms = ms < cand_ms ? ms : cand_ms;
                }
                if (rc) continue;
                if (cand_ms < best_ms) { best_ms = cand_ms; plan = cands[i].plan; }
            }
            // weight-ring granularity of the winner: three 18 KB slots (3 taps of a 96-filter use) instead of two 24 KB ones were
            // measured 2 % faster on KB block 2 and slower elsewhere (profiles/r02b_kb_ring_sweep.txt) -- let the clock decide
            for (int variant = 0; variant < 5; ++variant) {   // 0, 1: slot size; 2-4: which ring takes the spare shared memory first
                V2Plan alt = plan;
                if (variant < 2) alt.wcap = variant == 0 ? 18 * 1024 : 12 * 1024;
                else alt.ring = variant - 1;
                int rc = launch_with_plan(p, alt, map_src, n_maps, map_cls, has_up, sms, stream, false, of_src, n_of);
                float alt_ms = 1e30f;
                for (int rep = 0; rep < 3 && !rc; ++rep) {
                    KB_CUDA(cudaEventRecord(e0, stream));
                    rc = launch_with_plan(p, alt, map_src, n_maps, map_cls, has_up, sms, stream, false, of_src, n_of);
                    KB_CUDA(cudaEventRecord(e1, stream));
                    if (rc) break;
                    KB_CUDA(cudaEventSynchronize(e1));
                    float ms = 0.f;
                    KB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                    alt_ms = ms < alt_ms ? ms : alt_ms;
                }
                if (!rc && alt_ms < 0.985f * best_ms) { best_ms = alt_ms; plan = alt; }
            }
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        if (const char* ov = getenv("KBNET_V2_PLAN")) {  // debugging aid: "R,TW,n_split,dual"
            int a, b, c, d;
            if (sscanf(ov, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) plan = {a, b, c, d, 0, 0};
        }
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        if (g_plans.size() < 4096) g_plans.push_back({key, plan});
    }
    return launch_with_plan(p, plan, map_src, n_maps, map_cls, has_up, sms, stream, true, of_src, n_of);
}

// same grouping as conv_umma.cu / conv_pack_weights: every source is cut into chunks of <= 8 channels, padded to an even count
int chunk_list(const int* src_channels, int n_src, int* src_of, int* c0_of, int* nch_of) {
    int g = 0;
    for (int s = 0; s < n_src; ++s)
        for (int c0 = 0; c0 < src_channels[s]; c0 += 8) {
            if (g >= 2 * V2_MAX_STAGES) return -1;
            src_of[g] = s; c0_of[g] = c0; nch_of[g] = src_channels[s] - c0 < 8 ? src_channels[s] - c0 : 8;
            ++g;
        }
    if (g & 1) {
        if (g >= 2 * V2_MAX_STAGES) return -1;
        src_of[g] = -1; c0_of[g] = 0; nch_of[g] = 0;
        ++g;
    }
    return g;
}

}  // namespace
void set_autotune(int on) { g_autotune.store(on ? 1 : 0); }
namespace {

bool v2_disabled() {
    static const bool off = getenv("KBNET_CONV_V1") != nullptr;
    return off;
}

}  // namespace

// ---- plain convolution ------------------------------------------------------------------------------------------------
bool conv_v2_supported(const ConvParams& c) {
    if (v2_disabled()) return false;
    if (!((c.ksize == 3 || c.ksize == 1) && (c.stride == 1 || c.stride == 2))) return false;
    if (c.n_src < 1 || c.n_src > 4) return false;   // fp32 sources use tensor maps 0-3
    const int Npad = (c.F + 15) & ~15;
    if (Npad > 1024) return false;
    for (int i = 0; i < c.n_src; ++i) {
        const ConvSrcDev& s = c.src[i];
        if (!tma_ok(s)) return false;
        if (s.mode == SRC_DIRECT) continue;
        if (s.mode == SRC_UPSAMPLE && c.ksize == 3 && c.stride == 1 && c.Hin == 2 * s.H && c.Win == 2 * s.W) continue;
        return false;
    }
    int chans[CONV_MAX_SRC], so[2 * V2_MAX_STAGES], co[2 * V2_MAX_STAGES], no[2 * V2_MAX_STAGES];
    for (int i = 0; i < c.n_src; ++i) chans[i] = c.src[i].C;
    return chunk_list(chans, c.n_src, so, co, no) > 0;
}

int conv_v2_launch(const ConvParams& c, const void* wpack, cudaStream_t stream) {
    KB_CHECK_ARG(wpack && ((uintptr_t)wpack & 15) == 0, "conv(v2): packed weights missing or misaligned");
    static thread_local V2Params q;  // large: keep it off the stack
    memset(&q, 0, sizeof(q));
    q.geo = c.stride == 1 ? (c.ksize == 3 ? G_S1K3 : G_S1K1) : (c.ksize == 3 ? G_S2K3 : G_S2K1);
    q.N = c.N; q.Hin = c.Hin; q.Win = c.Win; q.Hout = c.Hout; q.Wout = c.Wout;
    q.slope = c.slope;
    q.n_convs = 1;
    V2Conv& cv = q.conv[0];
    cv.wpack = (const uint8_t*)wpack; cv.out = c.out; cv.out_stride_n = c.out_stride_n;
    cv.F = c.F; cv.Npad = (c.F + 15) & ~15; cv.act = c.act; cv.ntaps = c.ksize * c.ksize;
    cv.oplane = (long long)c.Hout * c.Wout; cv.o_off = 0; cv.o_row_pitch = c.Wout; cv.o_col_mul = 1;
    int chans[CONV_MAX_SRC], so[2 * V2_MAX_STAGES], co[2 * V2_MAX_STAGES], no[2 * V2_MAX_STAGES];
    for (int i = 0; i < c.n_src; ++i) chans[i] = c.src[i].C;
    const int ng = chunk_list(chans, c.n_src, so, co, no);
    KB_CHECK_ARG(ng > 0, "conv(v2): too many input channels");
    q.n_stages = ng / 2;
    bool has_up = false;
    uint8_t map_cls[V2_MAX_MAPS] = {0, 0, 0, 0};
    for (int i = 0; i < c.n_src; ++i) { map_cls[i] = c.src[i].mode == SRC_UPSAMPLE ? 1 : 0; has_up = has_up || map_cls[i]; }
    for (int s = 0; s < q.n_stages; ++s) {
        V2Stage& st = q.stage[s];
        for (int h = 0; h < 2; ++h) {
            const int g = 2 * s + h;
            V2Chunk& ch = st.ch[h];
            if (so[g] < 0) { ch.kind = CK_NONE; continue; }
            ch.kind = CK_TMA; ch.tmap = (uint8_t)so[g]; ch.nch = (uint8_t)no[g]; ch.c0 = (uint16_t)co[g]; ch.cls = map_cls[so[g]];
        }
        st.n_use = 1;
        st.use[0].conv = 0; st.use[0].tapmask = cv.ntaps == 9 ? TAPS_ALL9 : TAPS_ONE; st.use[0].first = s == 0; st.use[0].wtile = (uint32_t)(s * cv.ntaps);
    }
    return v2_launch(q, c.src, c.n_src, map_cls, has_up, stream);
}

// ---- the fused Calibrated Backprojection block ----------------------------------------------------------------------------
// One launch for reference src/net_utils.py:1343-1371.  `coordinates` (the [N,3,H,W] tensor the reference passes) or `kinv`.
bool kb_fused_supported(const float* image, long long image_sn, const float* depth, long long depth_sn, const float* fused, long long fused_sn,
                        const float* coordinates, int C_i, int C_d, int C_f, int H, int W, int F_i, int F_d, int F_f) {
    if (v2_disabled()) return false;
    if ((C_i & 15) || (C_d & 15) || (C_f & 7) || C_d > V2_MAX_ZC) return false;
    if ((W & 3) || W < 4) return false;
    auto okp = [](const float* p, long long sn) { return ((uintptr_t)p & 15) == 0 && (sn & 3) == 0; };
    if (!okp(image, image_sn) || !okp(depth, depth_sn) || (fused && !okp(fused, fused_sn)) || (coordinates && !okp(coordinates, 0))) return false;
    const int np_i = (F_i + 15) & ~15, np_d = (F_d + 15) & ~15, np_f = (F_f + 15) & ~15;
    // some power-of-two filter split must bring the three accumulators under 512 TMEM columns
    for (int ns = 1; ns <= 8; ns *= 2)
        if (np_i % (16 * ns) == 0 && np_d % (16 * ns) == 0 && np_f % (16 * ns) == 0 && (np_i + np_d + np_f) / ns <= 512 && np_i / ns <= 256 &&
            np_f / ns <= 256)
            return C_i / 16 + C_d / 16 + 1 + (C_f / 8 + 2) / 2 <= V2_MAX_STAGES;
    return false;
}

int kb_fused_launch(const float* image, long long image_sn, const float* depth, long long depth_sn, const float* fused, long long fused_sn,
                    const float* coordinates, const float* kinv, int N, int C_i, int C_d, int C_f, int H, int W, const float* w_proj,
                    const void* wpack_image, int F_i, const void* wpack_depth, int F_d, const void* wpack_fused, int F_f, float slope,
                    float* out_image, long long out_image_sn, float* out_depth, long long out_depth_sn, float* out_fused, long long out_fused_sn,
                    cudaStream_t stream, const KbOperandFormat* of) {
    static thread_local V2Params q;
    memset(&q, 0, sizeof(q));
    q.geo = G_S2K3;
    q.N = N; q.Hin = H; q.Win = W; q.Hout = (H - 1) / 2 + 1; q.Wout = (W - 1) / 2 + 1;
    q.slope = slope; q.kinv = kinv; q.w_proj = w_proj; q.C_z = C_d;
    const bool image_of = of && of->image_hi && of->image_lo;                 // image / fused arrive in operand format: TMA straight into
    const bool fused_of = of && fused && of->fused_hi && of->fused_lo;        // the operand planes, no converter work for their chunks
    const bool out_image_of = of && of->out_image_hi && of->out_image_lo;
    // conv_image with no output at all is not computed (KBNetEncoder's last block: src/networks.py:476-523 only reads its shape);
    // the image is then needed at even pixels only (conv_fused's centre tap)
    const bool skip_image = out_image == nullptr && !out_image_of;
    const int ci = skip_image ? -1 : 0, cd = skip_image ? 0 : 1, cf = skip_image ? 1 : 2;   // convolution slots
    q.n_convs = skip_image ? 2 : 3;
    auto set_conv = [&](int j, const void* pack, float* out, long long osn, int F, int ntaps) {
        V2Conv& cv = q.conv[j];
        cv.wpack = (const uint8_t*)pack; cv.out = out; cv.out_stride_n = osn;
        cv.F = F; cv.Npad = (F + 15) & ~15; cv.act = 1; cv.ntaps = ntaps;
        cv.oplane = (long long)q.Hout * q.Wout; cv.o_off = 0; cv.o_row_pitch = q.Wout; cv.o_col_mul = 1;
    };
    if (!skip_image) set_conv(ci, wpack_image, out_image, out_image_sn, F_i, 9);
    set_conv(cd, wpack_depth, out_depth, out_depth_sn, F_d, 9);
    set_conv(cf, wpack_fused, out_fused, out_fused_sn, F_f, 1);
    if (of) {   // operand-format copies of conv_image / conv_fused for the next block
        KB_CHECK_ARG((!out_image_of || F_i % 8 == 0) && (!(of->out_fused_hi && of->out_fused_lo) || F_f % 8 == 0),
                     "kb(v2): operand-format outputs need filter counts that are multiples of 8");
        if (out_image_of) { q.conv[ci].of_hi = (uint8_t*)of->out_image_hi; q.conv[ci].of_lo = (uint8_t*)of->out_image_lo; q.conv[ci].of_cg = F_i / 8; }
        if (of->out_fused_hi && of->out_fused_lo) {   // conv_fused: its even pixels only (all the next block reads)
            q.conv[cf].of_hi = (uint8_t*)of->out_fused_hi; q.conv[cf].of_lo = (uint8_t*)of->out_fused_lo; q.conv[cf].of_cg = F_f / 8; q.conv[cf].of_even = 1;
        }
        KB_CHECK_ARG(!(image_of || fused_of) || (!(H & 1) && !(W & 1)), "kb(v2): operand-format sources need even H and W");
        KB_CHECK_ARG(!(out_image_of || q.conv[cf].of_hi) || (!(q.Hout & 1) && !(q.Wout & 1)), "kb(v2): operand-format outputs need even Ho and Wo");
    }
    KB_CHECK_ARG(out_depth != nullptr && out_fused != nullptr, "kb(v2): conv_depth / conv_fused need their NCHW outputs");
    // tensor maps: 0 image, 1 depth, 2 fused, 3 coordinates (fp32 NCHW); 4/5 image, 6/7 fused in operand format
    ConvSrcDev srcs[V2_MAX_MAPS];
    memset(srcs, 0, sizeof(srcs));
    auto mk = [&](const float* ptr, int C, long long sn) {
        ConvSrcDev s;
        memset(&s, 0, sizeof(s));
        s.ptr = ptr; s.C = C; s.H = H; s.W = W; s.stride_n = sn;
        return s;
    };
    srcs[1] = mk(depth, C_d, depth_sn);
    srcs[0] = image_of ? srcs[1] : mk(image, C_i, image_sn);               // unused maps still have to encode: point them at depth
    srcs[2] = (fused && !fused_of) ? mk(fused, C_f, fused_sn) : srcs[1];
    srcs[3] = coordinates ? mk(coordinates, 3, 3LL * H * W) : srcs[1];
    // sources read at even pixels only (`fused`; the image when conv_image is skipped) use even-row boxes: half the traffic
    const uint8_t map_cls[V2_MAX_MAPS] = {(uint8_t)(skip_image ? 1 : 0), 0, 1, 0};
    OfSrc ofs[2];
    memset(ofs, 0, sizeof(ofs));
    int n_of = 0;
    if (image_of || fused_of) {
        n_of = 2;
        const OfSrc img = {of->image_hi, of->image_lo, 4 * (C_i / 8), H / 2, W / 2}, fus = {of->fused_hi, of->fused_lo, C_f / 8, H / 2, W / 2};
        ofs[0] = image_of ? img : fus;       // both map pairs must encode: an unused one aliases the other
        ofs[1] = fused_of ? fus : img;
    }
    int s = 0;
    // (1) depth chunks -> conv_depth (9 taps) and the z projection
    for (int j = 0; j < C_d / 16; ++j, ++s) {
        V2Stage& st = q.stage[s];
        for (int h = 0; h < 2; ++h) { st.ch[h].kind = CK_TMA; st.ch[h].tmap = 1; st.ch[h].nch = 8; st.ch[h].c0 = (uint16_t)(16 * j + 8 * h); st.ch[h].zacc = 1; }
        st.n_use = 1;
        st.use[0].conv = (uint8_t)cd; st.use[0].tapmask = TAPS_ALL9; st.use[0].first = j == 0; st.use[0].wtile = (uint32_t)(j * 9);
    }
    // (2) camera coordinates (3 channels + a zero chunk) -> conv_depth
    {
        V2Stage& st = q.stage[s];
        if (coordinates) { st.ch[0].kind = CK_TMA; st.ch[0].tmap = 3; st.ch[0].nch = 3; st.ch[0].c0 = 0; }
        else st.ch[0].kind = CK_COORDS;
        st.ch[1].kind = CK_NONE;
        st.n_use = 1;
        st.use[0].conv = (uint8_t)cd; st.use[0].tapmask = TAPS_ALL9; st.use[0].first = 0; st.use[0].wtile = (uint32_t)((C_d / 16) * 9);
        ++s;
    }
    // (3) image chunks -> conv_image (9 taps) and conv_fused (centre tap)
    for (int j = 0; j < C_i / 16; ++j, ++s) {
        V2Stage& st = q.stage[s];
        for (int h = 0; h < 2; ++h) {
            V2Chunk& ch = st.ch[h];
            ch.nch = 8; ch.c0 = (uint16_t)(16 * j + 8 * h);
            if (skip_image) { ch.kind = image_of ? CK_OF_P0 : CK_TMA_P0; ch.tmap = image_of ? 4 : 0; ch.cls = image_of ? 0 : 1; }
            else { ch.kind = image_of ? CK_OF : CK_TMA; ch.tmap = image_of ? 4 : 0; }
        }
        int u = 0;
        if (!skip_image) { st.use[u].conv = (uint8_t)ci; st.use[u].tapmask = TAPS_ALL9; st.use[u].first = j == 0; st.use[u].wtile = (uint32_t)(j * 9); ++u; }
        st.use[u].conv = (uint8_t)cf; st.use[u].tapmask = TAPS_CENTER; st.use[u].first = j == 0; st.use[u].wtile = (uint32_t)j; ++u;
        st.n_use = u;
    }
    // (4) xyz = coordinates * z, then the incoming fused features -> conv_fused (even pixels only)
    {
        const int n_rest = 1 + C_f / 8;  // chunks of conv_fused after the image: xyz, fused...
        for (int g = 0; g < n_rest + (n_rest & 1); g += 2, ++s) {
            V2Stage& st = q.stage[s];
            for (int h = 0; h < 2; ++h) {
                const int k = g + h;
                V2Chunk& ch = st.ch[h];
                if (k == 0) {
                    if (coordinates) { ch.kind = CK_XYZ_T; ch.tmap = 3; ch.nch = 3; ch.c0 = 0; }
                    else ch.kind = CK_XYZ;
                } else if (k < n_rest) {
                    if (fused_of) { ch.kind = CK_OF_P0; ch.tmap = (uint8_t)(image_of ? 6 : 4); ch.nch = 8; ch.c0 = (uint16_t)(8 * (k - 1)); ch.cls = 1; }
                    else { ch.kind = CK_TMA_P0; ch.tmap = 2; ch.nch = 8; ch.c0 = (uint16_t)(8 * (k - 1)); ch.cls = 1; }
                } else {
                    ch.kind = CK_NONE;
                }
            }
            st.n_use = 1;
            st.use[0].conv = (uint8_t)cf; st.use[0].tapmask = TAPS_CENTER; st.use[0].first = 0; st.use[0].wtile = (uint32_t)(C_i / 16 + g / 2);
        }
    }
    KB_CHECK_ARG(s <= V2_MAX_STAGES, "kb(v2): too many stages");
    q.n_stages = s;
    return v2_launch(q, srcs, 4, map_cls, false, stream, n_of ? ofs : nullptr, n_of);
}

// ---- nearest-2x upsample + 3x3 convolution as four 2x2 convolutions at the SOURCE resolution ---------------------------------
// (UpConv2d / DecoderBlock.deconv of the reference, src/net_utils.py:484-499,1472-1480.)  For an output pixel (2y+py, 2x+px) the
// three rows 2y+py-1 .. 2y+py+1 of the upsampled image are only TWO source rows -- {y-1, y, y} for py = 0, {y, y, y+1} for
// py = 1 -- and likewise for the columns, so each of the four output parities is a 2x2 convolution of the source with row/column
// sums of the 3x3 weights: 4 taps instead of 9 (2.25x fewer MMAs), on a grid of a quarter of the pixels, and the upsampled tensor
// is neither materialised nor even gathered.  Zero padding carries over exactly (source row -1 / H_s <=> upsampled row -1 / 2 H_s).
__global__ void upconv2x_weights_kernel(const float* __restrict__ w, int F, int C, float* __restrict__ out) {
    // out[parity = py*2+px][f][c][a][b], a/b = source row/column offset index (py + a - 1, px + b - 1 relative to (y, x))
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * C) return;
    float k[3][3];
    for (int j = 0; j < 9; ++j) k[j / 3][j % 3] = w[(size_t)i * 9 + j];
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            float r[2][3];  // rows combined
            for (int x = 0; x < 3; ++x) {
                r[0][x] = py == 0 ? k[0][x] : k[0][x] + k[1][x];
                r[1][x] = py == 0 ? k[1][x] + k[2][x] : k[2][x];
            }
            float* o = out + ((size_t)(py * 2 + px) * F * C + i) * 4;
            for (int a = 0; a < 2; ++a) {
                o[a * 2 + 0] = px == 0 ? r[a][0] : r[a][0] + r[a][1];
                o[a * 2 + 1] = px == 0 ? r[a][1] + r[a][2] : r[a][2];
            }
        }
}

int upconv2x_prepare_weights(const float* w, int F, int C, float* out4, cudaStream_t stream) {
    KB_CHECK_ARG(w && out4 && F > 0 && C > 0, "upconv: bad argument");
    upconv2x_weights_kernel<<<(F * C + 127) / 128, 128, 0, stream>>>(w, F, C, out4);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

bool upconv2x_supported(const ConvSrcDev& src, int F) {
    if (v2_disabled()) return false;
    const int Npad = (F + 15) & ~15;
    if (!tma_ok(src) || 4 * Npad > 2048 || Npad > 256) return false;
    return (src.C + 7) / 8 <= 2 * V2_MAX_STAGES;
}

// wpack4: four buffers of conv_pack_weights(w_parity[j], {C}, 1, F, ksize = 2); out is [N, F, 2H, 2W]
int upconv2x_launch(const ConvSrcDev& src, int N, const void* const* wpack4, int F, int act, float slope, float* out, long long out_stride_n,
                    cudaStream_t stream) {
    static thread_local V2Params q;
    memset(&q, 0, sizeof(q));
    q.geo = G_S1K3;
    q.N = N; q.Hin = src.H; q.Win = src.W; q.Hout = src.H; q.Wout = src.W;  // the launch grid is the SOURCE grid
    q.slope = slope;
    q.n_convs = 4;
    const int Wo = 2 * src.W;
    for (int j = 0; j < 4; ++j) {
        KB_CHECK_ARG(wpack4[j] && ((uintptr_t)wpack4[j] & 15) == 0, "upconv: packed weights missing or misaligned");
        V2Conv& cv = q.conv[j];
        const int py = j >> 1, px = j & 1;
        cv.wpack = (const uint8_t*)wpack4[j]; cv.out = out; cv.out_stride_n = out_stride_n;
        cv.F = F; cv.Npad = (F + 15) & ~15; cv.act = act; cv.ntaps = 4;
        cv.oplane = 4LL * src.H * src.W; cv.o_off = (long long)py * Wo + px; cv.o_row_pitch = 2 * Wo; cv.o_col_mul = 2;
    }
    int chans[1] = {src.C}, so[2 * V2_MAX_STAGES], co[2 * V2_MAX_STAGES], no[2 * V2_MAX_STAGES];
    const int ng = chunk_list(chans, 1, so, co, no);
    KB_CHECK_ARG(ng > 0, "upconv: too many input channels");
    q.n_stages = ng / 2;
    for (int s = 0; s < q.n_stages; ++s) {
        V2Stage& st = q.stage[s];
        for (int h = 0; h < 2; ++h) {
            const int g = 2 * s + h;
            V2Chunk& ch = st.ch[h];
            if (so[g] < 0) { ch.kind = CK_NONE; continue; }
            ch.kind = CK_TMA; ch.tmap = 0; ch.nch = (uint8_t)no[g]; ch.c0 = (uint16_t)co[g]; ch.cls = 0;
        }
        st.n_use = 4;
        for (int j = 0; j < 4; ++j) {
            const int py = j >> 1, px = j & 1;
            uint16_t mask = 0;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) mask |= (uint16_t)(1u << ((py + a) * 3 + (px + b)));  // source row y + py + a - 1, column x + px + b - 1
            st.use[j].conv = (uint8_t)j; st.use[j].tapmask = mask; st.use[j].first = s == 0; st.use[j].wtile = (uint32_t)(s * 4);
        }
    }
    const uint8_t map_cls[V2_MAX_MAPS] = {0, 0, 0, 0};
    return v2_launch(q, &src, 1, map_cls, false, stream);
}

}  // namespace kbnet
