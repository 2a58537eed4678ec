// Tensor-core convolution engine for sm_100a: im2col-free implicit GEMM on tcgen05 with TMEM accumulators.
//
//   out[n, f, oy, ox] = act( sum_{c, ky, kx} w[f, c, ky, kx] * in[n, c, oy*s - pad + ky, ox*s - pad + kx] )
//
// where `in` is the channel concatenation of up to 4 gathered sources (plain NCHW tensor, nearest-upsampled tensor,
// K^-1-generated camera coordinates, xyz = coordinates * z) -- never materialised.
//
// GEMM view per CTA tile (R x TW output pixels of one sample, all F filters):
//   M = output pixels, linearised q = oyl * P + oxl with pitch P = TW + halo, cut into MB blocks of 128 rows (TMEM lanes);
//   N = filters (padded to 16; split in parts of <= 256);  K = 16 input channels per stage, times the k*k taps.
// Operands are bf16 with the 3-term split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (fp32 accumulate in TMEM): 16 mantissa
// bits per operand, which keeps every layer inside the 1e-3 parity budget (single-pass bf16/tf32 does not).
//
// A operand (activations): producer warps gather fp32 NCHW from global, split to bf16 hi/lo and store K-major,
// NO-swizzle "chunk planes":  plane[cp][entry] = 16 bytes = 8 channels of one input pixel.  Because consecutive entries
// are consecutive pixels, the (ky,kx) tap of a 3x3 conv is the SAME plane read from a start address shifted by
// (ky*P + kx) * 16 bytes -- a descriptor change, no data movement (verified by tools/umma_probe.cu).  Stride-2 convs keep
// four parity planes (even/odd row x even/odd col) so that the same trick works.
// B operand (weights): pre-packed once per weight tensor to the exact shared-memory image (pack_weights_kernel) and
// streamed per (stage, tap) with 1-D TMA bulk copies through an mbarrier ring.
//
// Warp roles: warps 0-7 gather/convert producers, then epilogue (tcgen05.ld -> LeakyReLU -> coalesced NCHW stores);
// warp 8 lane 0 issues tcgen05.mma; warp 9 lane 0 runs the weight TMA ring.
#include "conv_common.cuh"
#include "umma.cuh"

namespace kbnet {

using namespace ptx;

constexpr int UM_MAX_GROUPS = 128;  // 8-channel groups -> up to 1024 (padded) input channels
constexpr int UM_PROD_WARPS = 8;
constexpr int UM_THREADS = (UM_PROD_WARPS + 2) * 32;
constexpr int UM_MAX_ASTAGES = 3;
constexpr int UM_MAX_WSLOTS = 8;
constexpr int UM_HEADER = 256;
constexpr int UM_SMEM_LIMIT = 227 * 1024;

enum { GEO_S1 = 0, GEO_S2K3 = 1, GEO_S2K1 = 2 };

struct UmmaConvParams {
    ConvSrcDev src[CONV_MAX_SRC];
    int n_src;
    int N, Hin, Win, Hout, Wout;
    int ksize, stride, geo, pad;
    int F, Npad, n_parts, part_N;
    int n_groups, n_stages;
    uint32_t groups[UM_MAX_GROUPS];  // src | nch << 8 | c0 << 16
    const uint8_t* wpack;            // [n_stages][taps][Npad * 64 B], tile layout in pack_weights_kernel
    float* out;
    long long out_stride_n;
    int act;
    float slope;
    int R, TW, P, MB, tiles_x, tiles_y;
    int n_planes, n_py, erows, n_entries, plane_px;
    int a_stages, w_slots;
    int tmem_cols;
};

// ---------------------------------------------------------------------------------------------------------------------
// group list shared by the packer and the kernel: every source is cut into groups of <= 8 channels, padded to an even count
// ---------------------------------------------------------------------------------------------------------------------
static int build_groups(const int* src_channels, int n_src, uint32_t* groups, int* n_groups) {
    int g = 0;
    for (int s = 0; s < n_src; ++s)
        for (int c0 = 0; c0 < src_channels[s]; c0 += 8) {
            if (g >= UM_MAX_GROUPS) return -1;
            const int nch = src_channels[s] - c0 < 8 ? src_channels[s] - c0 : 8;
            groups[g++] = (uint32_t)s | ((uint32_t)nch << 8) | ((uint32_t)c0 << 16);
        }
    if (g & 1) {
        if (g >= UM_MAX_GROUPS) return -1;
        groups[g++] = 0;  // zero group (nch = 0)
    }
    *n_groups = g;
    return 0;
}

struct PackParams {
    const float* w;
    uint8_t* out;
    int F, Ctot, taps, Npad, n_stages;
    int seg_off[CONV_MAX_SRC];
    uint32_t groups[UM_MAX_GROUPS];
};

__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
    const long long total = (long long)p.n_stages * p.taps * 2 * p.Npad * 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int e = (int)(idx & 7);
    long long r = idx >> 3;
    const int n = (int)(r % p.Npad); r /= p.Npad;
    const int c = (int)(r & 1); r >>= 1;
    const int t = (int)(r % p.taps);
    const int s = (int)(r / p.taps);
    const uint32_t g = p.groups[2 * s + c];
    const int src = g & 0xFF, nch = (g >> 8) & 0xFF, c0 = g >> 16;
    float v = 0.f;
    if (n < p.F && e < nch) v = __ldg(p.w + ((size_t)n * p.Ctot + p.seg_off[src] + c0 + e) * p.taps + t);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    // tile = [g = n / 8][hi | lo][chunk 0 | 1][8 rows][16 B]: any range of filter groups is contiguous (filter splits load with ONE
    // bulk copy), and every operand form is a K-major no-swizzle descriptor: LBO = 128 B between the K chunks, SBO = 512 B between
    // 8-filter groups (hi or lo rows alone), SBO = 256 B for the interleaved [hi | lo] rows of the "dual" MMA.
    uint8_t* tile = p.out + ((size_t)s * p.taps + t) * ((size_t)p.Npad * 64);
    const size_t off = (size_t)(n >> 3) * 512 + (size_t)c * 128 + (size_t)(n & 7) * 16 + e * 2;
    *reinterpret_cast<__nv_bfloat16*>(tile + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(tile + off + 256) = lo;
}

size_t conv_wpack_bytes(const int* src_channels, int n_src, int F, int ksize) {
    uint32_t groups[UM_MAX_GROUPS];
    int ng = 0;
    if (build_groups(src_channels, n_src, groups, &ng)) return 0;
    const int Npad = (F + 15) & ~15;
    return (size_t)(ng / 2) * ksize * ksize * Npad * 64;
}

int conv_pack_weights(const float* w, const int* src_channels, int n_src, int F, int ksize, void* wpack, cudaStream_t stream) {
    KB_CHECK_ARG(w && wpack && src_channels, "pack: null pointer argument");
    KB_CHECK_ARG(n_src >= 1 && n_src <= CONV_MAX_SRC && F > 0 && ksize >= 1 && ksize <= 3, "pack: bad argument");
    PackParams p;
    memset(&p, 0, sizeof(p));
    int ng = 0;
    KB_CHECK_ARG(build_groups(src_channels, n_src, p.groups, &ng) == 0, "pack: more than %d input channels", UM_MAX_GROUPS * 8);
    p.w = w; p.out = (uint8_t*)wpack; p.F = F; p.taps = ksize * ksize; p.Npad = (F + 15) & ~15; p.n_stages = ng / 2;
    p.Ctot = 0;
    for (int s = 0; s < n_src; ++s) { p.seg_off[s] = p.Ctot; p.Ctot += src_channels[s]; }
    const long long total = (long long)p.n_stages * p.taps * 2 * p.Npad * 8;
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------------------
// 8 channels [c0, c0+nch) of one input pixel (y, x) of source S; zeros outside the image (conv zero padding) and for c >= nch
__device__ __forceinline__ void gather8(const ConvSrcDev& S, int n, int c0, int nch, int y, int x, bool inside, float (&v)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    if (!inside) return;
    if (S.mode == SRC_DIRECT || S.mode == SRC_UPSAMPLE) {
        int sy = y, sx = x;
        if (S.mode == SRC_UPSAMPLE) { sy = nearest_src(y, S.scale_y, S.H); sx = nearest_src(x, S.scale_x, S.W); }
        const size_t plane = (size_t)S.H * S.W;
        const float* ptr = S.ptr + (size_t)n * S.stride_n + (size_t)c0 * plane + (size_t)sy * S.W + sx;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < nch) v[c] = __ldg(ptr + (size_t)c * plane);
    } else {
        float zz = 1.f;
        if (S.mode != SRC_COORDS_KINV) zz = __ldg(S.aux + (size_t)n * S.zplane + (size_t)(y >> 1) * S.zW + (x >> 1));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c0 + c < 3 && c < nch) {
                float co;
                if (S.mode == SRC_XYZ_TENSOR) {
                    co = __ldg(S.ptr + (size_t)n * S.stride_n + ((size_t)(c0 + c) * S.H + y) * S.W + x);
                } else {
                    const float* k = S.ptr + (size_t)n * 9 + 3 * (c0 + c);
                    co = fmaf(__ldg(k), (float)x, fmaf(__ldg(k + 1), (float)y, __ldg(k + 2)));
                }
                v[c] = co * zz;
            }
        }
    }
}

__device__ __forceinline__ void store_split8(uint8_t* hi_dst, uint8_t* lo_dst, const float (&v)[8]) {
    const uint2 a = split_bf16x2(v[0], v[1]), b = split_bf16x2(v[2], v[3]);
    const uint2 c = split_bf16x2(v[4], v[5]), d = split_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(a.x, b.x, c.x, d.x);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(a.y, b.y, c.y, d.y);
}

__global__ void __launch_bounds__(UM_THREADS, 2) conv_umma_kernel(const __grid_constant__ UmmaConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    // barrier indices
    const int B_AFULL = 0, B_AEMPTY = UM_MAX_ASTAGES, B_WFULL = 2 * UM_MAX_ASTAGES, B_WEMPTY = B_WFULL + UM_MAX_WSLOTS,
              B_ACC = B_WEMPTY + UM_MAX_WSLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (B_ACC + 1));
    const uint32_t a_stage_bytes = (uint32_t)p.n_planes * 4u * (uint32_t)p.plane_px * 16u;
    const uint32_t w_slot_bytes = (uint32_t)p.Npad * 64u;
    uint8_t* sA = smem + UM_HEADER;
    uint8_t* sW = sA + (size_t)p.a_stages * a_stage_bytes;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int b = blockIdx.x;
    const int tx = b % p.tiles_x; b /= p.tiles_x;
    const int ty = b % p.tiles_y;
    const int n = b / p.tiles_y;
    const int oy0 = ty * p.R, ox0 = tx * p.TW;
    const int taps = p.ksize * p.ksize;

    if (tid == 0) {
        for (int i = 0; i < UM_MAX_ASTAGES; ++i) { mbar_init(BAR(B_AFULL + i), UM_PROD_WARPS * 32); mbar_init(BAR(B_AEMPTY + i), 1); }
        for (int i = 0; i < UM_MAX_WSLOTS; ++i) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
        mbar_init(BAR(B_ACC), 1);
        fence_mbar_init();
    }
    if (warp == UM_PROD_WARPS) {
        tmem_alloc_dyn(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < UM_PROD_WARPS) {
        // =========================== producers: gather -> bf16 hi/lo -> chunk planes ===========================
        const int nblk = (p.n_entries + 31) >> 5;
        const int items = 2 * p.n_py * nblk;
        for (int s = 0; s < p.n_stages; ++s) {
            const int slot = s % p.a_stages;
            mbar_wait(BAR(B_AEMPTY + slot), (((uint32_t)(s / p.a_stages)) & 1u) ^ 1u);
            uint8_t* stage = sA + (size_t)slot * a_stage_bytes;
            // Loads of IPT items are issued back to back before any of them is converted: the gather is latency-bound,
            // so the number of independent global loads in flight per thread is what sets the fill time of a stage.
            if (p.geo == GEO_S2K3) {
                constexpr int IPT = 2;
                for (int base = warp; base < items; base += UM_PROD_WARPS * IPT) {
                    float v[IPT][2][8];
                    int ee[IPT], pyy[IPT], gii[IPT];
#pragma unroll
                    for (int u = 0; u < IPT; ++u) {
                        const int it = base + u * UM_PROD_WARPS;
                        ee[u] = -1;
                        if (it < items) {
                            const int blk = it % nblk, rest = it / nblk;
                            const int py = rest % p.n_py, gi = rest / p.n_py;
                            const int e = (blk << 5) + lane;
                            if (e < p.n_entries) {
                                const uint32_t g = p.groups[2 * s + gi];
                                const int nch = (g >> 8) & 0xFF, c0 = g >> 16;
                                const ConvSrcDev& S = p.src[g & 0xFF];
                                const int ei = e / p.P, ej = e - ei * p.P;
                                const int y = 2 * (oy0 - 1 + ei) + py;
                                const int x = 2 * (ox0 - 1 + ej);
                                const bool yin = nch > 0 && y >= 0 && y < p.Hin;
                                gather8(S, n, c0, nch, y, x, yin && x >= 0 && x < p.Win, v[u][0]);
                                gather8(S, n, c0, nch, y, x + 1, yin && x + 1 >= 0 && x + 1 < p.Win, v[u][1]);
                                ee[u] = e; pyy[u] = py; gii[u] = gi;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < IPT; ++u) {
                        if (ee[u] >= 0) {
#pragma unroll
                            for (int px = 0; px < 2; ++px) {
                                uint8_t* pl = stage + (size_t)((pyy[u] * 2 + px) * 4) * p.plane_px * 16;
                                store_split8(pl + ((size_t)gii[u] * p.plane_px + ee[u]) * 16,
                                             pl + ((size_t)(2 + gii[u]) * p.plane_px + ee[u]) * 16, v[u][px]);
                            }
                        }
                    }
                }
            } else {
                constexpr int IPT = 4;
                for (int base = warp; base < items; base += UM_PROD_WARPS * IPT) {
                    float v[IPT][8];
                    int ee[IPT], gii[IPT];
#pragma unroll
                    for (int u = 0; u < IPT; ++u) {
                        const int it = base + u * UM_PROD_WARPS;
                        ee[u] = -1;
                        if (it < items) {
                            const int blk = it % nblk, gi = it / nblk;
                            const int e = (blk << 5) + lane;
                            if (e < p.n_entries) {
                                const uint32_t g = p.groups[2 * s + gi];
                                const int nch = (g >> 8) & 0xFF, c0 = g >> 16;
                                const ConvSrcDev& S = p.src[g & 0xFF];
                                const int ei = e / p.P, ej = e - ei * p.P;
                                int y, x;
                                if (p.geo == GEO_S1) { y = oy0 - p.pad + ei; x = ox0 - p.pad + ej; }
                                else                 { y = 2 * (oy0 + ei);   x = 2 * (ox0 + ej); }
                                gather8(S, n, c0, nch, y, x, nch > 0 && y >= 0 && y < p.Hin && x >= 0 && x < p.Win, v[u]);
                                ee[u] = e; gii[u] = gi;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < IPT; ++u)
                        if (ee[u] >= 0)
                            store_split8(stage + ((size_t)gii[u] * p.plane_px + ee[u]) * 16,
                                         stage + ((size_t)(2 + gii[u]) * p.plane_px + ee[u]) * 16, v[u]);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(BAR(B_AFULL + slot));
        }
        // =========================== epilogue: TMEM -> registers -> activation -> NCHW ===========================
        mbar_wait(BAR(B_ACC), 0);
        tc_fence_after();
        const int quarter = warp & 3, half = warp >> 2;
        const int ncol_half = p.Npad >> 1;
        const size_t oplane = (size_t)p.Hout * p.Wout;
        for (int mb = 0; mb < p.MB; ++mb) {
            const int q = mb * 128 + quarter * 32 + lane;
            const int oyl = q / p.P, oxl = q - oyl * p.P;
            const int oy = oy0 + oyl, ox = ox0 + oxl;
            const bool valid = oyl < p.R && oxl < p.TW && oy < p.Hout && ox < p.Wout;
            float* dst = p.out + (size_t)n * p.out_stride_n + (size_t)oy * p.Wout + ox;
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * p.Npad + half * ncol_half);
            for (int c = 0; c < ncol_half; c += 8) {
                uint32_t r[8];
                tmem_ld_x8(trow + c, r);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int f = half * ncol_half + c + j;
                        if (f < p.F) {
                            float o = __uint_as_float(r[j]);
                            if (p.act) o = leaky(o, p.slope);
                            dst[(size_t)f * oplane] = o;
                        }
                    }
                }
            }
        }
    } else if (warp == UM_PROD_WARPS) {
        // =========================== MMA issuer (one thread) ===========================
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16_f32(128, p.part_N);
            const uint64_t adesc0 = smem_desc(0, (uint32_t)p.plane_px * 16u, 128u);
            const uint64_t bdesc0 = smem_desc(0, 128u, 512u);  // packed tile layout: see pack_weights_kernel
            const uint32_t a_lo_off = (2u * (uint32_t)p.plane_px * 16u) >> 4;
            const uint32_t b_lo_off = 256u >> 4;
            uint32_t wp = 0;
            for (int s = 0; s < p.n_stages; ++s) {
                const int slot = s % p.a_stages;
                mbar_wait(BAR(B_AFULL + slot), ((uint32_t)(s / p.a_stages)) & 1u);
                tc_fence_after();
                const uint32_t a_base = smem_u32(sA + (size_t)slot * a_stage_bytes);
                for (int t = 0; t < taps; ++t, ++wp) {
                    const int ws = wp % p.w_slots;
                    mbar_wait(BAR(B_WFULL + ws), (wp / p.w_slots) & 1u);
                    tc_fence_after();
                    int pl = 0, shift = 0;
                    if (p.geo == GEO_S1 && p.ksize == 3) {
                        shift = (t / 3) * p.P + (t % 3);
                    } else if (p.geo == GEO_S2K3) {
                        const int dy = t / 3 - 1, dx = t % 3 - 1;
                        pl = (dy != 0 ? 2 : 0) + (dx != 0 ? 1 : 0);
                        shift = (dy == -1 ? 0 : 1) * p.P + (dx == -1 ? 0 : 1);
                    }
                    const uint32_t a_hi = (a_base + (uint32_t)(pl * 4 * p.plane_px + shift) * 16u) >> 4;
                    const uint32_t w_hi = smem_u32(sW + (size_t)ws * w_slot_bytes) >> 4;
                    for (int mb = 0; mb < p.MB; ++mb) {
                        const uint64_t ad_hi = adesc0 + (uint64_t)(a_hi + (uint32_t)mb * 128u);
                        const uint64_t ad_lo = ad_hi + a_lo_off;
                        for (int pi = 0; pi < p.n_parts; ++pi) {
                            const uint64_t bd_hi = bdesc0 + (uint64_t)(w_hi + (uint32_t)(pi * p.part_N) * 4u);
                            const uint64_t bd_lo = bd_hi + b_lo_off;
                            const uint32_t d = tmem_base + (uint32_t)(mb * p.Npad + pi * p.part_N);
                            mma_bf16_ss(d, ad_hi, bd_hi, idesc, (s | t) ? 1u : 0u);
                            mma_bf16_ss(d, ad_lo, bd_hi, idesc, 1u);
                            mma_bf16_ss(d, ad_hi, bd_lo, idesc, 1u);
                        }
                    }
                    mma_commit(BAR(B_WEMPTY + ws));
                }
                mma_commit(BAR(B_AEMPTY + slot));
            }
            mma_commit(BAR(B_ACC));
        }
        __syncwarp();
    } else {
        // =========================== weight TMA ring (one thread) ===========================
        if (lane == 0) {
            const uint32_t npk = (uint32_t)p.n_stages * (uint32_t)taps;
            for (uint32_t pk = 0; pk < npk; ++pk) {
                const int ws = pk % p.w_slots;
                mbar_wait(BAR(B_WEMPTY + ws), ((pk / p.w_slots) & 1u) ^ 1u);
                mbar_arrive_expect_tx(BAR(B_WFULL + ws), w_slot_bytes);
                bulk_g2s(smem_u32(sW + (size_t)ws * w_slot_bytes), p.wpack + (size_t)pk * w_slot_bytes, w_slot_bytes, BAR(B_WFULL + ws));
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == UM_PROD_WARPS) tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tile selection + launch
// ---------------------------------------------------------------------------------------------------------------------
static double mma_cycles(int N) {  // measured, profiles/r01_umma_probe.md
    double c = N / 2.0, s = 32.0 + N / 4.0;
    c = c > s ? c : s;
    return c > 47.0 ? c : 47.0;
}

static int roundup(int a, int b) { return (a + b - 1) / b * b; }

bool conv_umma_supported(const ConvParams& c) {
    if (!((c.ksize == 3 && (c.stride == 1 || c.stride == 2)) || (c.ksize == 1 && (c.stride == 1 || c.stride == 2)))) return false;
    int chans[CONV_MAX_SRC];
    for (int i = 0; i < c.n_src; ++i) chans[i] = c.src[i].C;
    uint32_t groups[UM_MAX_GROUPS];
    int ng;
    if (build_groups(chans, c.n_src, groups, &ng)) return false;
    const int Npad = (c.F + 15) & ~15;
    return Npad <= 512;
}

int conv_umma_launch(const ConvParams& c, const void* wpack, cudaStream_t stream) {
    KB_CHECK_ARG(wpack, "conv(umma): packed weights missing");
    KB_CHECK_ARG(conv_umma_supported(c), "conv(umma): unsupported configuration (k=%d stride=%d F=%d)", c.ksize, c.stride, c.F);
    KB_CHECK_ARG(((uintptr_t)wpack & 15) == 0, "conv(umma): packed weights must be 16-byte aligned");
    UmmaConvParams p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < c.n_src; ++i) p.src[i] = c.src[i];
    p.n_src = c.n_src;
    p.N = c.N; p.Hin = c.Hin; p.Win = c.Win; p.Hout = c.Hout; p.Wout = c.Wout;
    p.ksize = c.ksize; p.stride = c.stride; p.pad = c.ksize / 2;
    p.geo = c.stride == 1 ? GEO_S1 : (c.ksize == 3 ? GEO_S2K3 : GEO_S2K1);
    p.F = c.F; p.Npad = (c.F + 15) & ~15;
    p.n_parts = (p.Npad + 255) / 256;
    while (p.Npad % p.n_parts || (p.Npad / p.n_parts) % 16) ++p.n_parts;
    p.part_N = p.Npad / p.n_parts;
    int chans[CONV_MAX_SRC];
    for (int i = 0; i < c.n_src; ++i) chans[i] = c.src[i].C;
    KB_CHECK_ARG(build_groups(chans, c.n_src, p.groups, &p.n_groups) == 0, "conv(umma): too many input channels");
    p.n_stages = p.n_groups / 2;
    p.wpack = (const uint8_t*)wpack;
    p.out = c.out; p.out_stride_n = c.out_stride_n; p.act = c.act; p.slope = c.slope;
    p.n_planes = p.geo == GEO_S2K3 ? 4 : 1;
    p.n_py = p.geo == GEO_S2K3 ? 2 : 1;
    const int taps = c.ksize * c.ksize;
    const int halo_x = p.geo == GEO_S1 ? 2 * p.pad : (p.geo == GEO_S2K3 ? 1 : 0);
    const int halo_y = halo_x;
    const int w_slot_bytes = p.Npad * 64;
    const int mb_max = 512 / p.Npad;
    // ---- tile search: minimise waves * per-tile cycles under the smem / TMEM limits ------------------------------
    // A tile that needs <= 256 TMEM columns and <= half of the shared memory lets two CTAs share an SM: one CTA's epilogue
    // and prologue then overlap the other's MMA main loop (the kernel itself is not persistent).
    constexpr int SMEM_HALF = 113 * 1024;
    auto plane_px_of = [&](int R, int P, int MB) {
        const int n_entries = (R + halo_y) * P;
        const int max_shift = p.geo == GEO_S1 ? (taps == 9 ? 2 * P + 2 : 0) : (p.geo == GEO_S2K3 ? P + 1 : 0);
        return roundup(n_entries > max_shift + MB * 128 ? n_entries : max_shift + MB * 128, 8);
    };
    double best = 1e30;
    int bR = 0, bTW = 0;
    for (int parts = 1; parts <= 16; ++parts) {
        const int TW = cdiv(c.Wout, parts);
        if (parts > 1 && TW == cdiv(c.Wout, parts - 1)) continue;
        const int P = TW + halo_x;
        for (int R = 1; R <= 64 && R <= c.Hout; ++R) {
            const int MB = cdiv(R * P, 128);
            if (MB > mb_max) break;
            const int n_entries = (R + halo_y) * P;
            const long long a_stage = (long long)p.n_planes * 4 * plane_px_of(R, P, MB) * 16;
            const long long smem_min = UM_HEADER + 2 * a_stage + 2LL * w_slot_bytes;
            if (smem_min > UM_SMEM_LIMIT) break;
            const bool co2 = MB * p.Npad <= 256 && smem_min <= SMEM_HALF;
            const long long tiles = (long long)c.N * cdiv(c.Hout, R) * cdiv(c.Wout, TW);
            const double mma = (double)p.n_stages * taps * MB * p.n_parts * 3.0 * mma_cycles(p.part_N);
            const double load = (double)p.n_stages * p.n_planes * n_entries * 64.0 / 20.0;
            const double wts = (double)p.n_stages * taps * w_slot_bytes / 40.0;
            const double epi = (double)MB * 128 * p.Npad * 4.0 / 20.0;
            double tile = mma > load ? mma : load;
            tile = (tile > wts ? tile : wts) + (co2 ? 0.25 : 1.0) * (epi + 6000.0);
            const double total = (double)cdiv((int)(tiles > (1 << 30) ? (1 << 30) : tiles), 148) * tile;
            if (total < best) { best = total; bR = R; bTW = TW; }
        }
    }
    KB_CHECK_ARG(bR > 0, "conv(umma): no tile fits (Npad=%d)", p.Npad);
    p.R = bR; p.TW = bTW; p.P = bTW + halo_x;
    p.MB = cdiv(p.R * p.P, 128);
    p.erows = p.R + halo_y;
    p.n_entries = p.erows * p.P;
    p.plane_px = plane_px_of(p.R, p.P, p.MB);
    const long long a_stage = (long long)p.n_planes * 4 * p.plane_px * 16;
    p.a_stages = 2;
    p.w_slots = 2;
    long long used = UM_HEADER + p.a_stages * a_stage + (long long)p.w_slots * w_slot_bytes;
    const long long budget = (p.MB * p.Npad <= 256 && used <= SMEM_HALF) ? SMEM_HALF : UM_SMEM_LIMIT;
    while (p.w_slots < UM_MAX_WSLOTS && p.w_slots < p.n_stages * taps && used + w_slot_bytes <= budget) { ++p.w_slots; used += w_slot_bytes; }
    if (p.n_stages > 2 && used + a_stage <= budget) { ++p.a_stages; used += a_stage; }
    p.tmem_cols = 32;
    while (p.tmem_cols < p.MB * p.Npad) p.tmem_cols <<= 1;
    p.tiles_x = cdiv(c.Wout, p.TW);
    p.tiles_y = cdiv(c.Hout, p.R);
    const long long grid = (long long)p.tiles_x * p.tiles_y * c.N;
    KB_CHECK_ARG(grid > 0 && grid < (1ll << 31), "conv(umma): grid too large");
    KB_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)used));
    conv_umma_kernel<<<(unsigned)grid, UM_THREADS, (size_t)used, stream>>>(p);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace kbnet
