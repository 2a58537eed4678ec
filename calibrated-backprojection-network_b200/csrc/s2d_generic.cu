// Generic (exact-fp32, CUDA-core) fused SparseToDensePool forward for sm_100a -- the fallback of s2d_fused.cu for
// configurations outside its fast path (input_channels > 2, very many pool sizes).  Same fusion, FFMA convolutions.
//
// Replaces reference src/networks.py:2168-2196 (~35 ATen launches, every intermediate plane through HBM)
// with ONE kernel: per output tile the sparse-depth halo tile is staged in shared memory once, all min/max
// pool radii are produced by a separable nested-window sweep held in registers, the 1x1 fusion chain runs
// in registers, and the 3x3 conv reads its 10-channel input from shared memory.  HBM traffic = read x once
// (+halo from L2), write out once.
//
// Tile geometry (template): TH x TW output pixels per CTA.
//   pyramid region : rows [y0-1, y0+TH+1)  cols [x0-4, x0+TW+4)      (PH x PW, 3x3-conv halo, 4-aligned in x)
//   z tile         : pyramid region grown by R (max pool radius) on every side (ZH x ZW)
// Pool semantics (bit-exact vs ATen, see oracle/pools_np.py): empty <=> z == 0; min over non-empty values with
// 999 as the "empty" flag and a result of exactly 999 mapped to 0; max with -inf padding; NaN propagates.
#include "common.cuh"

namespace kbnet {

constexpr int S2D_NF = 8;          // n_filter supported by the register-resident 1x1 chain
constexpr int S2D_MAX_POOL = 16;
constexpr int S2D_MAX_CONV = 8;
constexpr int S2D_MAX_CIN = 8;

struct S2DGenericParams {
    const float* x;
    float* out;
    float* pyramid;
    int N, C_in, H, W;
    int n_pool, n_min;
    int radius[S2D_MAX_POOL];
    unsigned min_mask, max_mask;  // bit r set <=> some min / max pool has radius r
    int rmax_min, rmax_max;
    const float* w1x1[S2D_MAX_CONV];
    int n_conv;
    const float* w3x3;
    float slope;
    int tiles_x, tiles_y;
};

template <int TH, int TW, int R, int LV, int LH, int NT>
__global__ void __launch_bounds__(NT) s2d_generic_kernel(const S2DGenericParams p) {
    constexpr int PH = TH + 2;
    constexpr int PW = TW + 8;
    constexpr int ZH = PH + 2 * R;
    constexpr int ZW = PW + 2 * R;
    constexpr int NF = S2D_NF;
    static_assert(PH % LV == 0, "vertical strips must tile the pyramid rows");
    static_assert(PW % LH == 0 && LH % 4 == 0, "horizontal strips are float4 multiples");
    static_assert(R % 4 == 0 && TW % 4 == 0, "alignment");

    extern __shared__ __align__(16) float smem[];
    const int n_cin_tot = NF + p.C_in;
    // layout: [w1x1 | w3x3 | V planes | z tiles (fmin,fmax)  -- later aliased by the conv input planes]
    const int n_w1 = NF * p.n_pool + (p.n_conv - 1) * NF * NF;
    const int n_w3 = n_cin_tot * 9 * NF;
    float* s_w1 = smem;
    float* s_w3 = s_w1 + ((n_w1 + 3) & ~3);
    float* s_V = s_w3 + ((n_w3 + 3) & ~3);
    float* s_zmin = s_V + p.n_pool * PH * ZW;
    float* s_zmax = s_zmin + ZH * ZW;
    float* s_cin = s_zmin;  // conv input planes [NF + C_in][PH][PW] alias the (dead) z tiles

    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const int tx = b % p.tiles_x;
    b /= p.tiles_x;
    const int ty = b % p.tiles_y;
    const int n = b / p.tiles_y;
    const int x0 = tx * TW, y0 = ty * TH;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;
    const float* xin = p.x + (size_t)n * p.C_in * plane;

    // ---- stage weights -------------------------------------------------------------------------
    for (int i = tid; i < NF * p.n_pool; i += NT) s_w1[i] = __ldg(p.w1x1[0] + i);
    for (int l = 1; l < p.n_conv; ++l)
        for (int i = tid; i < NF * NF; i += NT) s_w1[NF * p.n_pool + (l - 1) * NF * NF + i] = __ldg(p.w1x1[l] + i);
    // w3x3 [j][c][dy][dx] -> smem [c][dy][dx][j] so the 8 output filters of one tap are one 32-byte row
    for (int i = tid; i < n_w3; i += NT) {
        const int j = i / (n_cin_tot * 9), rem = i % (n_cin_tot * 9);
        s_w3[rem * NF + j] = __ldg(p.w3x3 + i);
    }

    // ---- 1. stage the z halo tile, already transformed for the two chains -----------------------
    {
        const int gy0 = y0 - 1 - R, gx0 = x0 - 4 - R;
        for (int idx = tid; idx < ZH * ZW; idx += NT) {
            const int i = idx / ZW, j = idx - i * ZW;
            const int gy = gy0 + i, gx = gx0 + j;
            float fmn = __int_as_float(0x7f800000), fmx = __int_as_float(0xff800000);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const float z = __ldg(xin + (size_t)gy * W + gx);
                fmn = (z == 0.f) ? 999.f : z;
                fmx = z;
            }
            s_zmin[idx] = fmn;
            s_zmax[idx] = fmx;
        }
    }
    __syncthreads();

    // ---- 2. vertical nested-window sweep: V[k][row][col] = min/max over rows [row-r_k, row+r_k] ---
    {
        constexpr int NSV = PH / LV;
        for (int s = tid; s < ZW * NSV; s += NT) {
            const int cx = s % ZW, js = s / ZW;
            const int rbase = js * LV;
            if (p.n_min > 0) {
                float f[LV + 2 * R], m[LV];
#pragma unroll
                for (int i = 0; i < LV + 2 * R; ++i) f[i] = s_zmin[(rbase + i) * ZW + cx];
#pragma unroll
                for (int i = 0; i < LV; ++i) m[i] = f[R + i];
#pragma unroll
                for (int r = 1; r <= R; ++r) {
                    if (r <= p.rmax_min) {
#pragma unroll
                        for (int i = 0; i < LV; ++i) m[i] = min_nan(m[i], min_nan(f[R + i - r], f[R + i + r]));
                        if ((p.min_mask >> r) & 1u) {
                            for (int k = 0; k < p.n_min; ++k)
                                if (p.radius[k] == r) {
                                    float* dst = s_V + (k * PH + rbase) * ZW + cx;
#pragma unroll
                                    for (int i = 0; i < LV; ++i) dst[i * ZW] = m[i];
                                }
                        }
                    }
                }
            }
            if (p.n_pool > p.n_min) {
                float f[LV + 2 * R], m[LV];
#pragma unroll
                for (int i = 0; i < LV + 2 * R; ++i) f[i] = s_zmax[(rbase + i) * ZW + cx];
#pragma unroll
                for (int i = 0; i < LV; ++i) m[i] = f[R + i];
#pragma unroll
                for (int r = 1; r <= R; ++r) {
                    if (r <= p.rmax_max) {
#pragma unroll
                        for (int i = 0; i < LV; ++i) m[i] = max_nan(m[i], max_nan(f[R + i - r], f[R + i + r]));
                        if ((p.max_mask >> r) & 1u) {
                            for (int k = p.n_min; k < p.n_pool; ++k)
                                if (p.radius[k] == r) {
                                    float* dst = s_V + (k * PH + rbase) * ZW + cx;
#pragma unroll
                                    for (int i = 0; i < LV; ++i) dst[i * ZW] = m[i];
                                }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- 3. horizontal sweep -> pyramid values in registers -> 1x1 fusion chain -> conv input planes
    {
        constexpr int NSH = PW / LH;
        const float slope = p.slope;
        for (int s = tid; s < PH * NSH; s += NT) {
            const int pr = s / NSH, hs = s - pr * NSH;
            const int c0 = hs * LH;
            const int gy = y0 - 1 + pr, gxb = x0 - 4 + c0;
            const bool row_in = (gy >= 0 && gy < H);
            const bool core_row = (pr >= 1 && pr <= TH);
            float acc[NF][LH];
#pragma unroll
            for (int j = 0; j < NF; ++j)
#pragma unroll
                for (int i = 0; i < LH; ++i) acc[j][i] = 0.f;

            for (int k = 0; k < p.n_pool; ++k) {
                const float* src = s_V + (k * PH + pr) * ZW + c0;
                float v[LH + 2 * R], m[LH];
#pragma unroll
                for (int q = 0; q < (LH + 2 * R) / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(src + 4 * q);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < LH; ++i) m[i] = v[R + i];
                const int rk = p.radius[k];
                if (k < p.n_min) {
#pragma unroll
                    for (int r = 1; r <= R; ++r)
                        if (r <= rk) {
#pragma unroll
                            for (int i = 0; i < LH; ++i) m[i] = min_nan(m[i], min_nan(v[R + i - r], v[R + i + r]));
                        }
#pragma unroll
                    for (int i = 0; i < LH; ++i) m[i] = (m[i] == 999.f) ? 0.f : m[i];
                } else {
#pragma unroll
                    for (int r = 1; r <= R; ++r)
                        if (r <= rk) {
#pragma unroll
                            for (int i = 0; i < LH; ++i) m[i] = max_nan(m[i], max_nan(v[R + i - r], v[R + i + r]));
                        }
                }
                if (p.pyramid != nullptr && row_in && core_row) {
                    float* dst = p.pyramid + ((size_t)n * p.n_pool + k) * plane + (size_t)gy * W;
#pragma unroll
                    for (int i = 0; i < LH; ++i) {
                        const int c = c0 + i, gx = gxb + i;
                        if (c >= 4 && c < TW + 4 && gx < W) dst[gx] = m[i];
                    }
                }
                const float* w0 = s_w1 + k;
#pragma unroll
                for (int j = 0; j < NF; ++j) {
                    const float w = w0[j * p.n_pool];
#pragma unroll
                    for (int i = 0; i < LH; ++i) acc[j][i] = fmaf(w, m[i], acc[j][i]);
                }
            }
#pragma unroll
            for (int j = 0; j < NF; ++j)
#pragma unroll
                for (int i = 0; i < LH; ++i) acc[j][i] = leaky(acc[j][i], slope);

            for (int l = 1; l < p.n_conv; ++l) {
                const float* wl = s_w1 + NF * p.n_pool + (l - 1) * NF * NF;
                float nxt[NF][LH];
#pragma unroll
                for (int j = 0; j < NF; ++j) {
                    const float4 wa = *reinterpret_cast<const float4*>(wl + j * NF);
                    const float4 wb = *reinterpret_cast<const float4*>(wl + j * NF + 4);
#pragma unroll
                    for (int i = 0; i < LH; ++i) {
                        float a = wa.x * acc[0][i];
                        a = fmaf(wa.y, acc[1][i], a);
                        a = fmaf(wa.z, acc[2][i], a);
                        a = fmaf(wa.w, acc[3][i], a);
                        a = fmaf(wb.x, acc[4][i], a);
                        a = fmaf(wb.y, acc[5][i], a);
                        a = fmaf(wb.z, acc[6][i], a);
                        a = fmaf(wb.w, acc[7][i], a);
                        nxt[j][i] = leaky(a, slope);
                    }
                }
#pragma unroll
                for (int j = 0; j < NF; ++j)
#pragma unroll
                    for (int i = 0; i < LH; ++i) acc[j][i] = nxt[j][i];
            }
            // zero padding of the 3x3 conv: pixels outside the image contribute 0
            bool inb[LH];
#pragma unroll
            for (int i = 0; i < LH; ++i) {
                const int gx = gxb + i;
                inb[i] = row_in && gx >= 0 && gx < W;
            }
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                float4 t;
                t.x = inb[0] ? acc[j][0] : 0.f;
                t.y = inb[1] ? acc[j][1] : 0.f;
                t.z = inb[2] ? acc[j][2] : 0.f;
                t.w = inb[3] ? acc[j][3] : 0.f;
                *reinterpret_cast<float4*>(s_cin + (j * PH + pr) * PW + c0) = t;
                if (LH == 8) {
                    float4 u;
                    u.x = inb[4 % LH] ? acc[j][4 % LH] : 0.f;
                    u.y = inb[5 % LH] ? acc[j][5 % LH] : 0.f;
                    u.z = inb[6 % LH] ? acc[j][6 % LH] : 0.f;
                    u.w = inb[7 % LH] ? acc[j][7 % LH] : 0.f;
                    *reinterpret_cast<float4*>(s_cin + (j * PH + pr) * PW + c0 + 4) = u;
                }
            }
            for (int c = 0; c < p.C_in; ++c) {
                const float* src = xin + (size_t)c * plane + (size_t)gy * W;
#pragma unroll
                for (int i = 0; i < LH; ++i)
                    s_cin[((NF + c) * PH + pr) * PW + c0 + i] = inb[i] ? __ldg(src + gxb + i) : 0.f;
            }
        }
    }
    __syncthreads();

    // ---- 4. 3x3 conv over [q(8), x(C_in)] + LeakyReLU, 4 pixels x 8 filters per thread -----------
    {
        constexpr int NSX = TW / 4;
        const bool vec_ok = (W % 4 == 0);
        for (int s = tid; s < TH * NSX; s += NT) {
            const int oy = s / NSX, ox = (s - oy * NSX) * 4;
            const int gy = y0 + oy, gx = x0 + ox;
            if (gy >= H || gx >= W) continue;
            float acc[NF][4];
#pragma unroll
            for (int j = 0; j < NF; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
            for (int c = 0; c < n_cin_tot; ++c) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const float* row = s_cin + (c * PH + oy + dy) * PW + ox + 3;  // pyramid col of (ox - 1)
                    float in[6];
                    in[0] = row[0];
                    const float4 t = *reinterpret_cast<const float4*>(row + 1);
                    in[1] = t.x; in[2] = t.y; in[3] = t.z; in[4] = t.w;
                    in[5] = row[5];
                    const float* wr = s_w3 + (c * 3 + dy) * 3 * NF;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float4 wa = *reinterpret_cast<const float4*>(wr + dx * NF);
                        const float4 wb = *reinterpret_cast<const float4*>(wr + dx * NF + 4);
                        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int j = 0; j < NF; ++j)
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wv[j], in[i + dx], acc[j][i]);
                    }
                }
            }
            float* dst = p.out + ((size_t)n * NF) * plane + (size_t)gy * W + gx;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                float4 t;
                t.x = leaky(acc[j][0], p.slope);
                t.y = leaky(acc[j][1], p.slope);
                t.z = leaky(acc[j][2], p.slope);
                t.w = leaky(acc[j][3], p.slope);
                float* d = dst + (size_t)j * plane;
                if (vec_ok) {
                    *reinterpret_cast<float4*>(d) = t;
                } else {
                    d[0] = t.x;
                    if (gx + 1 < W) d[1] = t.y;
                    if (gx + 2 < W) d[2] = t.z;
                    if (gx + 3 < W) d[3] = t.w;
                }
            }
        }
    }
}

template <int TH, int TW, int R, int LV, int LH, int NT>
static int launch_s2d_generic(S2DGenericParams& p, cudaStream_t stream) {
    constexpr int PH = TH + 2, PW = TW + 8, ZH = PH + 2 * R, ZW = PW + 2 * R;
    const int n_cin_tot = S2D_NF + p.C_in;
    const int n_w1 = S2D_NF * p.n_pool + (p.n_conv - 1) * S2D_NF * S2D_NF;
    const int n_w3 = n_cin_tot * 9 * S2D_NF;
    const size_t tail = (size_t)2 * ZH * ZW > (size_t)n_cin_tot * PH * PW ? (size_t)2 * ZH * ZW : (size_t)n_cin_tot * PH * PW;
    const size_t smem = sizeof(float) * (((n_w1 + 3) & ~3) + ((n_w3 + 3) & ~3) + (size_t)p.n_pool * PH * ZW + tail);
    KB_CHECK_ARG(smem <= 227 * 1024, "s2d: configuration needs %zu bytes of shared memory (> 227 KB)", smem);
    auto kern = s2d_generic_kernel<TH, TW, R, LV, LH, NT>;
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.tiles_x = cdiv(p.W, TW);
    p.tiles_y = cdiv(p.H, TH);
    const long long grid = (long long)p.tiles_x * p.tiles_y * p.N;
    KB_CHECK_ARG(grid > 0 && grid < (1ll << 31), "s2d: grid too large");
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int s2d_generic_forward(const float* x, int N, int C_in, int H, int W, const int* min_sizes, int n_min, const int* max_sizes,
                        int n_max, const float* const* w_1x1, int n_conv, int n_filter, const float* w_3x3, float neg_slope,
                        float* out, float* pyramid, cudaStream_t stream) {
    if (N == 0) return 0;
    KB_CHECK_ARG(x && out && w_3x3 && w_1x1, "s2d: null pointer argument");
    KB_CHECK_ARG(N >= 0 && H > 0 && W > 0, "s2d: bad shape N=%d H=%d W=%d", N, H, W);
    KB_CHECK_ARG(C_in >= 1 && C_in <= S2D_MAX_CIN, "s2d: input_channels=%d unsupported (1..%d)", C_in, S2D_MAX_CIN);
    KB_CHECK_ARG(n_filter == S2D_NF, "s2d: n_filter=%d unsupported (only %d)", n_filter, S2D_NF);
    KB_CHECK_ARG(n_conv >= 1 && n_conv <= S2D_MAX_CONV, "s2d: n_convolution=%d unsupported (1..%d)", n_conv, S2D_MAX_CONV);
    KB_CHECK_ARG(n_min >= 0 && n_max >= 0 && n_min + n_max >= 1 && n_min + n_max <= S2D_MAX_POOL,
                 "s2d: %d pool sizes unsupported (1..%d)", n_min + n_max, S2D_MAX_POOL);
    if (N == 0) return 0;
    S2DGenericParams p;
    p.x = x; p.out = out; p.pyramid = pyramid;
    p.N = N; p.C_in = C_in; p.H = H; p.W = W;
    p.n_pool = n_min + n_max; p.n_min = n_min;
    p.min_mask = p.max_mask = 0; p.rmax_min = p.rmax_max = 0;
    for (int k = 0; k < S2D_MAX_POOL; ++k) p.radius[k] = 0;
    for (int k = 0; k < p.n_pool; ++k) {
        const int s = k < n_min ? min_sizes[k] : max_sizes[k - n_min];
        KB_CHECK_ARG(s >= 3 && (s & 1) && s <= 33, "s2d: pool size %d unsupported (odd, 3..33)", s);
        const int r = s / 2;
        p.radius[k] = r;
        if (k < n_min) { p.min_mask |= 1u << r; if (r > p.rmax_min) p.rmax_min = r; }
        else           { p.max_mask |= 1u << r; if (r > p.rmax_max) p.rmax_max = r; }
    }
    for (int l = 0; l < S2D_MAX_CONV; ++l) p.w1x1[l] = l < n_conv ? w_1x1[l] : nullptr;
    for (int l = 0; l < n_conv; ++l) KB_CHECK_ARG(p.w1x1[l], "s2d: w_1x1[%d] is null", l);
    p.n_conv = n_conv; p.w3x3 = w_3x3; p.slope = neg_slope;
    const int rmax = p.rmax_min > p.rmax_max ? p.rmax_min : p.rmax_max;
    if (rmax <= 4) return launch_s2d_generic<16, 64, 4, 9, 4, 256>(p, stream);
    if (rmax <= 8) return launch_s2d_generic<16, 64, 8, 9, 4, 256>(p, stream);
    return launch_s2d_generic<16, 64, 16, 9, 4, 256>(p, stream);
}

}  // namespace kbnet
