#include "kernels_bwd.cuh"
#include "gemm_tc.cuh"

namespace dxmi {

namespace {
struct alignas(16) bf16x8 {
    __nv_bfloat162 v[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
// one LDG.128 / STG.128 (a member-wise struct copy compiles to four 4-byte accesses, see kernels.cu)
__device__ __forceinline__ bf16x8 ld8(const bf16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    bf16x8 r;
    *reinterpret_cast<uint4*>(&r) = u;
    return r;
}
__device__ __forceinline__ void st8(bf16* p, const bf16x8& v) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v); }
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
    bf16x8 p;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return p;
}
}  // namespace

// ================================================================================================ value head
// one CTA per image; thread = 8-channel vector x pixel lane
__global__ void __launch_bounds__(256) value_head_bwd_k(const bf16* __restrict__ h, const float* __restrict__ dout,
                                                       const float* __restrict__ lin_w, const float* __restrict__ scale_w,
                                                       bf16* __restrict__ dz, float* __restrict__ S, int HW, int C) {
    extern __shared__ float sred[];  // [PL][C]
    const int n = blockIdx.x;
    const int CV = C / 8, PL = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const float g = dout[n] * (scale_w ? scale_w[0] : 1.f);
    float s[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j] = 0.f;
        w[j] = g * lin_w[cv * 8 + j];
    }
    if (pl < PL) {
        for (int p = pl; p < HW; p += PL) {
            const long long off = ((long long)n * HW + p) * C + cv * 8;
            float f[8], d[8];
            unpack8(ld8(h + off), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += fmaxf(f[j], 0.f);
                d[j] = f[j] > 0.f ? w[j] : 0.f;
            }
            st8(dz + off, pack8(d));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sred[pl * C + cv * 8 + j] = s[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f;
        for (int q = 0; q < PL; ++q) a += sred[q * C + c];
        S[(long long)n * C + c] = a;
    }
}
void value_head_bwd(const bf16* h, const float* dout, const float* lin_w, const float* scale_w, bf16* dz, float* S, int N, int HW,
                    int C, cudaStream_t st) {
    const int PL = 256 / (C / 8);
    value_head_bwd_k<<<N, 256, (size_t)PL * C * sizeof(float), st>>>(h, dout, lin_w, scale_w, dz, S, HW, C);
}

// one CTA; thread c loops over the images in order
__global__ void __launch_bounds__(256) value_head_param_grads_k(const float* __restrict__ S, const float* __restrict__ dout,
                                                               const float* __restrict__ lin_w, const float* __restrict__ lin_b,
                                                               const float* __restrict__ scale_w, float* g_lin_w, float* g_lin_b,
                                                               float* g_scale_w, float* g_scale_b, int N, int C) {
    __shared__ float spre[1024];
    const float sw = scale_w ? scale_w[0] : 1.f;
    // pre[n] = S[n,:] . lin_w + lin_b  (one warp per image, fixed lane order)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < N; n += 8) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(S[(long long)n * C + c], lin_w[c], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) spre[n] = a + lin_b[0];
    }
    __syncthreads();
    // d lin_w: 4 independent partial sums per channel (images n = j mod 4), combined in a fixed order
    for (int c = threadIdx.x; c < C; c += 256) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        int n = 0;
        for (; n + 4 <= N; n += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = fmaf(dout[n + j] * sw, S[(long long)(n + j) * C + c], a[j]);
        }
        for (; n < N; ++n) a[0] = fmaf(dout[n] * sw, S[(long long)n * C + c], a[0]);
        if (g_lin_w) g_lin_w[c] = (a[0] + a[1]) + (a[2] + a[3]);
    }
    if (threadIdx.x == 0) {
        float gb = 0.f, gsw = 0.f, gsb = 0.f;
        for (int n = 0; n < N; ++n) {
            gb += dout[n] * sw;
            gsw = fmaf(dout[n], spre[n], gsw);
            gsb += dout[n];
        }
        if (g_lin_b) g_lin_b[0] = gb;
        if (g_scale_w) g_scale_w[0] = gsw;
        if (g_scale_b) g_scale_b[0] = gsb;
    }
}
void value_head_param_grads(const float* S, const float* dout, const float* lin_w, const float* lin_b, const float* scale_w,
                            float* g_lin_w, float* g_lin_b, float* g_scale_w, float* g_scale_b, int N, int C, cudaStream_t st) {
    value_head_param_grads_k<<<1, 256, 0, st>>>(S, dout, lin_w, lin_b, scale_w, g_lin_w, g_lin_b, g_scale_w, g_scale_b, N, C);
}

// ================================================================================================ pooling
__global__ void avgpool2_bwd_k(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int H, int W, int CV) {
    const long long total = (long long)N * H * W * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int x = (int)(r % W);
        r /= W;
        const int y = (int)(r % H);
        const int n = (int)(r / H);
        float f[8];
        unpack8(ld8(dy + ((((long long)n * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * CV + cv) * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= 0.25f;
        st8(dx + i * 8, pack8(f));
    }
}
void avgpool2_bwd(const bf16* dy, bf16* dx, int N, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)N * H * W * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 2368) blocks = 2368;
    avgpool2_bwd_k<<<(int)blocks, 256, 0, st>>>(dy, dx, N, H, W, C / 8);
}

// ================================================================================================ column sums
static int colsum_ctas(long long rows) {
    long long c = (rows + 255) / 256;
    if (c > 296) c = 296;
    if (c < 1) c = 1;
    return (int)c;
}
long long colsum_ws_floats(long long rows, int C) { return (long long)colsum_ctas(rows) * C; }

__global__ void __launch_bounds__(256) colsum_bf16_k(const bf16* __restrict__ x, long long rows, int C, float* __restrict__ ws) {
    extern __shared__ float sred[];  // [PL][C]
    const int CV = C / 8, PL = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const long long per = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = blockIdx.x * per;
    long long r1 = r0 + per;
    if (r1 > rows) r1 = rows;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (pl < PL) {
        long long r = r0 + pl;
        for (; r + 3LL * PL < r1; r += 4LL * PL) {  // 4 independent 16-byte loads in flight per thread
            bf16x8 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld8(x + (r + (long long)u * PL) * C + cv * 8);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += f[j];
            }
        }
        for (; r < r1; r += PL) {
            float f[8];
            unpack8(ld8(x + r * C + cv * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sred[pl * C + cv * 8 + j] = s[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f;
        for (int q = 0; q < PL; ++q) a += sred[q * C + c];
        ws[(long long)blockIdx.x * C + c] = a;
    }
}
// out[m] = sum_r ws[r][m].  Block = 32 columns x 8 row lanes: row lane j sums rows j, j+8, ... (coalesced 128-byte reads),
// then the 8 lanes are combined through smem in a fixed order (deterministic).
__global__ void __launch_bounds__(256) reduce_rows_f32_k(const float* __restrict__ ws, int R, int M, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int col = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int m = blockIdx.x * 32 + col;
    float a = 0.f;
    if (m < M)
        for (int r = rl; r < R; r += 8) a += ws[(long long)r * M + m];
    red[rl][col] = a;
    __syncthreads();
    if (rl == 0 && m < M) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += red[j][col];
        out[m] = s;
    }
}
void colsum_bf16(const bf16* x, long long rows, int C, float* ws, float* out, cudaStream_t st) {
    const int ctas = colsum_ctas(rows);
    const int PL = 256 / (C / 8);
    colsum_bf16_k<<<ctas, 256, (size_t)PL * C * sizeof(float), st>>>(x, rows, C, ws);
    reduce_rows_f32_k<<<(C + 31) / 32, 256, 0, st>>>(ws, ctas, C, out);
}

// ================================================================================================ first conv wgrad
// one CTA per (image, slab of FW_ROWS rows), one thread per output channel: 27 running sums over the slab's pixels; the
// zero-padded x patch of the slab sits in smem
static constexpr int FW_ROWS = 4;
__global__ void conv_first_wgrad_k(const bf16* __restrict__ dz, const float* __restrict__ x, float* __restrict__ ws, int H, int W,
                                   int Cout) {
    extern __shared__ float sx[];  // [3][FW_ROWS+2][W+2]
    const int n = blockIdx.x, h0 = blockIdx.y * FW_ROWS;
    const int pw = W + 2, ph = FW_ROWS + 2;
    for (int i = threadIdx.x; i < 3 * ph * pw; i += blockDim.x) {
        const int ci = i / (ph * pw), r = (i / pw) % ph, c = i % pw;
        const int hh = h0 + r - 1, ww = c - 1;
        sx[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((long long)n * 3 + ci) * H + hh) * W + ww] : 0.f;
    }
    __syncthreads();
    const int co = threadIdx.x;
    if (co >= Cout) return;
    float acc[27];
#pragma unroll
    for (int j = 0; j < 27; ++j) acc[j] = 0.f;
    const bf16* dp = dz + ((long long)n * H + h0) * W * Cout + co;
    for (int h = 0; h < FW_ROWS; ++h) {
        for (int w = 0; w < W; ++w) {
            const float d = __bfloat162float(dp[(long long)(h * W + w) * Cout]);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 3; ++q) acc[ci * 9 + r * 3 + q] = fmaf(d, sx[(ci * ph + h + r) * pw + w + q], acc[ci * 9 + r * 3 + q]);
        }
    }
#pragma unroll
    for (int j = 0; j < 27; ++j) ws[(((long long)n * gridDim.y + blockIdx.y) * Cout + co) * 27 + j] = acc[j];
}
void conv_first_wgrad(const bf16* dz, const float* x, float* ws, float* grad, int N, int H, int W, int Cout, cudaStream_t st) {
    const size_t smem = (size_t)3 * (FW_ROWS + 2) * (W + 2) * sizeof(float);
    const int threads = (Cout + 31) / 32 * 32;
    dim3 grid(N, H / FW_ROWS);
    conv_first_wgrad_k<<<grid, threads, smem, st>>>(dz, x, ws, H, W, Cout);
    const int M = Cout * 27;
    reduce_rows_f32_k<<<(M + 31) / 32, 256, 0, st>>>(ws, N * (H / FW_ROWS), M, grad);
}

// ================================================================================================ GroupNorm backward
static constexpr int GNB_THREADS = 256;
static int gnb_slabs(int HW, int C) {
    // a function of the geometry only (fixed summation order for every batch size)
    const int PL = GNB_THREADS / (C / 8) > 0 ? GNB_THREADS / (C / 8) : 1;
    int slabs = 1;
    while (slabs < 16 && HW / (slabs * 2) >= 4 * PL && HW % (slabs * 2) == 0) slabs *= 2;
    return slabs;
}
long long gn_bwd_ws_floats(int N, int HW, int C) { return (long long)N * gnb_slabs(HW, C) * C * 2 + (long long)N * C * 2; }

__device__ __forceinline__ float silu_grad(float z) {
    const float s = __fdividef(1.f, 1.f + __expf(-z));
    return s * fmaf(z, 1.f - s, 1.f);
}
__device__ __forceinline__ const bf16* gnb_src(const bf16* x1, int C1, const bf16* x2, int C2, long long pix, int c) {
    return (c < C1) ? x1 + pix * C1 + c : x2 + pix * C2 + (c - C1);
}

// pass 1: partial[(n * slabs + slab)][c] = (sum dz, sum dz * xh) over the slab's pixels
__global__ void __launch_bounds__(GNB_THREADS) gn_bwd_stats_k(const bf16* __restrict__ x1, int C1, const bf16* __restrict__ x2, int C2,
                                                              const bf16* __restrict__ dy, const float2* __restrict__ ab,
                                                              const float2* __restrict__ mr, int HW, int groups, int silu, int slabs,
                                                              float2* __restrict__ partial) {
    extern __shared__ float2 sred2[];  // [PL][C]
    const int C = C1 + C2, CV = C / 8, PL = GNB_THREADS / CV, cpg = C / groups;
    const int n = blockIdx.x, slab = blockIdx.y, pps = HW / slabs;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    if (pl < PL) {
        float a[8], b[8], mu[8], rs[8], sA[8], sB[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            const float2 v = ab[(long long)n * C + c], m = mr[(long long)n * groups + c / cpg];
            a[j] = v.x; b[j] = v.y; mu[j] = m.x; rs[j] = m.y;
            sA[j] = sB[j] = 0.f;
        }
        const long long base = (long long)n * HW + (long long)slab * pps;
        for (int p = pl; p < pps; p += PL) {
            float xf[8], df[8];
            unpack8(ld8(gnb_src(x1, C1, x2, C2, base + p, cv * 8)), xf);
            unpack8(ld8(dy + (base + p) * C + cv * 8), df);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dz = silu ? df[j] * silu_grad(fmaf(a[j], xf[j], b[j])) : df[j];
                sA[j] += dz;
                sB[j] = fmaf(dz, (xf[j] - mu[j]) * rs[j], sB[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sred2[pl * C + cv * 8 + j] = make_float2(sA[j], sB[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GNB_THREADS) {
        float2 t = make_float2(0.f, 0.f);
        for (int q = 0; q < PL; ++q) {
            const float2 v = sred2[q * C + c];
            t.x += v.x;
            t.y += v.y;
        }
        partial[((long long)n * slabs + slab) * C + c] = t;
    }
}

// pass 2: per-image totals -> group sums -> dx; the slab-0 CTA also publishes AB[n][c] for the parameter gradients
__global__ void __launch_bounds__(GNB_THREADS) gn_bwd_apply_k(const bf16* __restrict__ x1, int C1, const bf16* __restrict__ x2, int C2,
                                                              const bf16* __restrict__ dy, const float2* __restrict__ ab,
                                                              const float2* __restrict__ mr, int HW, int groups, int silu, int slabs,
                                                              const float2* __restrict__ partial, float2* __restrict__ AB,
                                                              bf16* __restrict__ dx) {
    __shared__ float2 s_ch[2048];
    __shared__ float s_SA[32], s_SB[32];
    const int C = C1 + C2, CV = C / 8, PL = GNB_THREADS / CV, cpg = C / groups;
    const int n = blockIdx.x, slab = blockIdx.y, pps = HW / slabs;
    for (int c = threadIdx.x; c < C; c += GNB_THREADS) {
        float2 t = make_float2(0.f, 0.f);
        for (int s = 0; s < slabs; ++s) {
            const float2 v = partial[((long long)n * slabs + s) * C + c];
            t.x += v.x;
            t.y += v.y;
        }
        if (slab == 0) AB[(long long)n * C + c] = t;
        // weight by gamma_c = a / rstd for the group sums
        const float gam = ab[(long long)n * C + c].x / mr[(long long)n * groups + c / cpg].y;
        s_ch[c] = make_float2(t.x * gam, t.y * gam);
    }
    __syncthreads();
    {
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        float s = 0.f, q = 0.f;
        if (g < groups)
            for (int i = sub; i < cpg; i += 8) {
                const float2 v = s_ch[g * cpg + i];
                s += v.x;
                q += v.y;
            }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && sub == 0) {
            const float inv_m = 1.f / ((float)cpg * (float)HW);
            s_SA[g] = s * inv_m;
            s_SB[g] = q * inv_m;
        }
    }
    __syncthreads();
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    if (pl >= PL) return;
    float a[8], b[8], mu[8], rs[8], sa[8], sb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j, g = c / cpg;
        const float2 v = ab[(long long)n * C + c], m = mr[(long long)n * groups + g];
        a[j] = v.x; b[j] = v.y; mu[j] = m.x; rs[j] = m.y;
        sa[j] = s_SA[g]; sb[j] = s_SB[g];
    }
    const long long base = (long long)n * HW + (long long)slab * pps;
    for (int p = pl; p < pps; p += PL) {
        float xf[8], df[8], o[8];
        unpack8(ld8(gnb_src(x1, C1, x2, C2, base + p, cv * 8)), xf);
        unpack8(ld8(dy + (base + p) * C + cv * 8), df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dz = silu ? df[j] * silu_grad(fmaf(a[j], xf[j], b[j])) : df[j];
            const float xh = (xf[j] - mu[j]) * rs[j];
            // rstd * (gamma dz - SA/m - xh SB/m) with a = rstd * gamma
            o[j] = fmaf(a[j], dz, -rs[j] * fmaf(xh, sb[j], sa[j]));
        }
        st8(dx + (base + p) * C + cv * 8, pack8(o));
    }
}

// dgamma[c] = sum_n AB[n][c].y, dbeta[c] = sum_n AB[n][c].x.  Block = 32 channels x 8 image lanes (lane j sums images j, j+8, ...),
// combined through smem in a fixed order.
__global__ void __launch_bounds__(256) gn_bwd_param_k(const float2* __restrict__ AB, int N, int C, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta) {
    __shared__ float2 red[8][33];
    const int col = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + col;
    float2 a = make_float2(0.f, 0.f);
    if (c < C)
        for (int n = rl; n < N; n += 8) {
            const float2 v = AB[(long long)n * C + c];
            a.x += v.x;
            a.y += v.y;
        }
    red[rl][col] = a;
    __syncthreads();
    if (rl == 0 && c < C) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sa += red[j][col].x;
            sb += red[j][col].y;
        }
        if (dbeta) dbeta[c] = sa;
        if (dgamma) dgamma[c] = sb;
    }
}

void group_norm_bwd(const bf16* x1, int C1, const bf16* x2, int C2, const bf16* dy, const float* ab, const float* mr, int N, int HW,
                    int groups, int silu, float* ws, bf16* dx, float* dgamma, float* dbeta, cudaStream_t st) {
    const int C = C1 + C2;
    const int slabs = gnb_slabs(HW, C);
    const int PL = GNB_THREADS / (C / 8) > 0 ? GNB_THREADS / (C / 8) : 1;
    float2* partial = reinterpret_cast<float2*>(ws);
    float2* AB = partial + (long long)N * slabs * C;
    dim3 grid(N, slabs);
    gn_bwd_stats_k<<<grid, GNB_THREADS, (size_t)PL * C * sizeof(float2), st>>>(x1, C1, x2, C2, dy, reinterpret_cast<const float2*>(ab),
                                                                              reinterpret_cast<const float2*>(mr), HW, groups, silu,
                                                                              slabs, partial);
    gn_bwd_apply_k<<<grid, GNB_THREADS, 0, st>>>(x1, C1, x2, C2, dy, reinterpret_cast<const float2*>(ab),
                                                 reinterpret_cast<const float2*>(mr), HW, groups, silu, slabs, partial, AB, dx);
    if (dgamma || dbeta) gn_bwd_param_k<<<(C + 31) / 32, 256, 0, st>>>(AB, N, C, dgamma, dbeta);
}

// ================================================================================================ U-Net backward helpers
__global__ void transpose_bf16_k(const bf16* __restrict__ src, long long ld_src, long long bs_src, bf16* __restrict__ dst, int R, int C) {
    __shared__ bf16 tile[32][33];
    const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const bf16* s = src + (long long)b * bs_src;
    bf16* d = dst + (long long)b * R * C;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? s[(long long)r * ld_src + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) d[(long long)c * R + r] = tile[threadIdx.x][i];
    }
}
void transpose_bf16_batched(const bf16* src, long long ld_src, long long bs_src, bf16* dst, int R, int C, int B, cudaStream_t st) {
    dim3 grid((C + 31) / 32, (R + 31) / 32, B), block(32, 8);
    transpose_bf16_k<<<grid, block, 0, st>>>(src, ld_src, bs_src, dst, R, C);
}

// one warp per row
__global__ void softmax_bwd_rows_k(const bf16* __restrict__ P, const float* __restrict__ dP, bf16* __restrict__ dS, long long rows, int S,
                                   float scale) {
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const bf16* p = P + row * S;
    const float* g = dP + row * S;
    float t = 0.f;
    for (int j = lane; j < S; j += 32) t = fmaf(__bfloat162float(p[j]), g[j], t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    for (int j = lane; j < S; j += 32) dS[row * S + j] = __float2bfloat16_rn(__bfloat162float(p[j]) * (g[j] - t) * scale);
}
void softmax_bwd_rows(const bf16* P, const float* dP, bf16* dS, long long rows, int S, float scale, cudaStream_t st) {
    const long long threads = rows * 32;
    softmax_bwd_rows_k<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, dP, dS, rows, S, scale);
}

__global__ void __launch_bounds__(256) attn_small_bwd_k(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, bf16* __restrict__ dqkv,
                                                       int S, int C, float scale) {
    extern __shared__ uint8_t smb[];
    bf16* sq = reinterpret_cast<bf16*>(smb);  // [S][C]
    bf16* sk = sq + S * C;
    bf16* sv = sk + S * C;
    bf16* sdo = sv + S * C;
    float* sp = reinterpret_cast<float*>(sdo + S * C);  // [S][S] probabilities
    float* sds = sp + S * S;                             // [S][S] dP, then dS
    const int n = blockIdx.x;
    const long long b3 = (long long)n * S * 3 * C, b1 = (long long)n * S * C;
    for (int i = threadIdx.x; i < S * C; i += 256) {
        const int t = i / C, c = i % C;
        sq[i] = qkv[b3 + (long long)t * 3 * C + c];
        sk[i] = qkv[b3 + (long long)t * 3 * C + C + c];
        sv[i] = qkv[b3 + (long long)t * 3 * C + 2 * C + c];
        sdo[i] = d_o[b1 + i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * S; i += 256) {
        const int a = i / S, b = i % S;
        float s = 0.f, g = 0.f;
        for (int c = 0; c < C; ++c) {
            s = fmaf(__bfloat162float(sq[a * C + c]), __bfloat162float(sk[b * C + c]), s);
            g = fmaf(__bfloat162float(sdo[a * C + c]), __bfloat162float(sv[b * C + c]), g);
        }
        sp[i] = s * scale;
        sds[i] = g;  // dP
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int a = threadIdx.x >> 5; a < S; a += 8) {
        float m = -INFINITY;
        for (int b = lane; b < S; b += 32) m = fmaxf(m, sp[a * S + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float l = 0.f;
        for (int b = lane; b < S; b += 32) {
            const float e = __expf(sp[a * S + b] - m);
            sp[a * S + b] = e;
            l += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        const float inv = 1.f / l;
        float t = 0.f;
        for (int b = lane; b < S; b += 32) {
            const float p = sp[a * S + b] * inv;
            sp[a * S + b] = p;
            t = fmaf(p, sds[a * S + b], t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        for (int b = lane; b < S; b += 32) sds[a * S + b] = sp[a * S + b] * (sds[a * S + b] - t) * scale;  // dS
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * C; i += 256) {
        const int t = i / C, c = i % C;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < S; ++j) {
            dq = fmaf(sds[t * S + j], __bfloat162float(sk[j * C + c]), dq);   // dQ[t] = sum_j dS[t,j] K[j]
            dk = fmaf(sds[j * S + t], __bfloat162float(sq[j * C + c]), dk);   // dK[t] = sum_i dS[i,t] Q[i]
            dv = fmaf(sp[j * S + t], __bfloat162float(sdo[j * C + c]), dv);   // dV[t] = sum_i P[i,t] dO[i]
        }
        dqkv[b3 + (long long)t * 3 * C + c] = __float2bfloat16_rn(dq);
        dqkv[b3 + (long long)t * 3 * C + C + c] = __float2bfloat16_rn(dk);
        dqkv[b3 + (long long)t * 3 * C + 2 * C + c] = __float2bfloat16_rn(dv);
    }
}
void attn_small_bwd(const bf16* qkv, const bf16* d_o, bf16* dqkv, int N, int S, int C, float scale, cudaStream_t st) {
    const size_t smem = (size_t)4 * S * C * sizeof(bf16) + (size_t)2 * S * S * sizeof(float);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(attn_small_bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured.set();
    }
    attn_small_bwd_k<<<N, 256, smem, st>>>(qkv, d_o, dqkv, S, C, scale);
}

__global__ void zero_insert2x_k(const bf16* __restrict__ dy, bf16* __restrict__ out, int N, int h, int w, int CV) {
    const long long total = (long long)N * (2 * h) * (2 * w) * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int x = (int)(r % (2 * w));
        r /= 2 * w;
        const int y = (int)(r % (2 * h));
        const int n = (int)(r / (2 * h));
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((x & 1) && (y & 1)) v = *reinterpret_cast<const uint4*>(dy + ((((long long)n * h + (y >> 1)) * w + (x >> 1)) * CV + cv) * 8);
        *reinterpret_cast<uint4*>(out + i * 8) = v;
    }
}
void zero_insert2x(const bf16* dy, bf16* out, int N, int h, int w, int C, cudaStream_t st) {
    const long long total = (long long)N * 4 * h * w * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 2368) blocks = 2368;
    zero_insert2x_k<<<(int)blocks, 256, 0, st>>>(dy, out, N, h, w, C / 8);
}

__global__ void sumpool2_k(const bf16* __restrict__ dy, bf16* __restrict__ out, int N, int h, int w, int CV) {
    const long long total = (long long)N * h * w * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int x = (int)(r % w);
        r /= w;
        const int y = (int)(r % h);
        const int n = (int)(r / h);
        const bf16* p = dy + ((((long long)n * 2 * h + 2 * y) * 2 * w + 2 * x) * CV + cv) * 8;
        float a[8], b[8], c[8], d[8], o[8];
        unpack8(ld8(p), a);
        unpack8(ld8(p + CV * 8), b);
        unpack8(ld8(p + (long long)2 * w * CV * 8), c);
        unpack8(ld8(p + (long long)2 * w * CV * 8 + CV * 8), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (a[j] + b[j]) + (c[j] + d[j]);
        st8(out + i * 8, pack8(o));
    }
}
void sumpool2(const bf16* dy, bf16* out, int N, int h, int w, int C, cudaStream_t st) {
    const long long total = (long long)N * h * w * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 2368) blocks = 2368;
    sumpool2_k<<<(int)blocks, 256, 0, st>>>(dy, out, N, h, w, C / 8);
}

__global__ void accum_bf16_k(bf16* __restrict__ dst, const bf16* __restrict__ src, long long ld_src, long long rows, int CV, int init) {
    const long long total = rows * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        const long long r = i / CV;
        float s[8];
        unpack8(ld8(src + r * ld_src + cv * 8), s);
        if (!init) {
            float d[8];
            unpack8(ld8(dst + i * 8), d);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += d[j];
        }
        st8(dst + i * 8, pack8(s));
    }
}
void accum_bf16(bf16* dst, const bf16* src, long long ld_src, long long rows, int C, int init, cudaStream_t st) {
    const long long total = rows * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 2368) blocks = 2368;
    accum_bf16_k<<<(int)blocks, 256, 0, st>>>(dst, src, ld_src, rows, C / 8, init);
}

// one CTA per image
__global__ void __launch_bounds__(256) colsum_per_image_k(const bf16* __restrict__ x, int HW, int C, float* __restrict__ out, int ld_out) {
    extern __shared__ float sred[];
    const int n = blockIdx.x;
    const int CV = C / 8, PL = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (pl < PL) {
        for (int p = pl; p < HW; p += PL) {
            float f[8];
            unpack8(ld8(x + ((long long)n * HW + p) * C + cv * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sred[pl * C + cv * 8 + j] = s[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f;
        for (int q = 0; q < PL; ++q) a += sred[q * C + c];
        out[(long long)n * ld_out + c] = a;
    }
}
void colsum_per_image(const bf16* x, int N, int HW, int C, float* out, int ld_out, cudaStream_t st) {
    const int PL = 256 / (C / 8);
    colsum_per_image_k<<<N, 256, (size_t)PL * C * sizeof(float), st>>>(x, HW, C, out, ld_out);
}

__device__ __forceinline__ float silu_exact_f(float v) { return v / (1.f + expf(-v)); }
__global__ void linear_bwd_w_k(const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, int ld_x, int act_x,
                               float* __restrict__ dW, float* __restrict__ db, int N, int O, int K) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)O * K) return;
    const int o = (int)(i / K), k = (int)(i % K);
    float a = 0.f, b = 0.f;
    for (int n = 0; n < N; ++n) {
        const float g = dy[(long long)n * ld_dy + o];
        float xv = x[(long long)n * ld_x + k];
        if (act_x == 2) xv = silu_exact_f(xv);
        a = fmaf(g, xv, a);
        b += g;
    }
    if (dW) dW[i] = a;
    if (db && k == 0) db[o] = b;
}
void linear_bwd_w(const float* dy, int ld_dy, const float* x, int ld_x, int act_x, float* dW, float* db, int N, int O, int K,
                  cudaStream_t st) {
    const long long total = (long long)O * K;
    linear_bwd_w_k<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, ld_dy, x, ld_x, act_x, dW, db, N, O, K);
}
__global__ void linear_bwd_x_k(const float* __restrict__ dy, int ld_dy, const float* __restrict__ W, float* __restrict__ dx, int ld_dx,
                               int N, int O, int K, int accumulate) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    float a = 0.f;
    for (int o = 0; o < O; ++o) a = fmaf(dy[(long long)n * ld_dy + o], W[(long long)o * K + k], a);
    float* d = dx + (long long)n * ld_dx + k;
    *d = accumulate ? *d + a : a;
}
void linear_bwd_x(const float* dy, int ld_dy, const float* W, float* dx, int ld_dx, int N, int O, int K, int accumulate, cudaStream_t st) {
    const long long total = (long long)N * K;
    linear_bwd_x_k<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, ld_dy, W, dx, ld_dx, N, O, K, accumulate);
}
// ---- the whole temb_proj stack in one launch each (24 tiny per-layer launches of ~60 us were 11 % of the U-Net training step)
__global__ void __launch_bounds__(256) linear_stack_bwd_w_k(const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, int ld_x,
                                                            const __grid_constant__ LinearStack ls, int N, int K) {
    // one CTA = 8 output rows (stacked row index) x 256 k; the 8 dy columns of a sample are broadcast loads
    __shared__ float s_dy[8][128];
    const int total = ls.off[ls.n_layers];
    const int r0 = blockIdx.x * 8;
    for (int n0 = 0; n0 < N; n0 += 128) {  // stage dy[n0 .. n0+127][r0 .. r0+7] (N <= 128 per pass)
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * 128; i += 256) {
            const int rr = i & 7, n = n0 + (i >> 3);
            s_dy[rr][i >> 3] = (n < N && r0 + rr < total) ? dy[(long long)n * ld_dy + r0 + rr] : 0.f;
        }
        __syncthreads();
        for (int k = blockIdx.y * 256 + threadIdx.x; k < K; k += gridDim.y * 256) {
            float acc[8], bs[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) acc[rr] = bs[rr] = 0.f;
            const int nn = N - n0 < 128 ? N - n0 : 128;
            for (int n = 0; n < nn; ++n) {
                const float xv = x[(long long)(n0 + n) * ld_x + k];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    acc[rr] = fmaf(s_dy[rr][n], xv, acc[rr]);
                    bs[rr] += s_dy[rr][n];
                }
            }
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int r = r0 + rr;
                if (r >= total) break;
                int l = 0;
                while (r >= ls.off[l + 1]) ++l;
                const int o = r - ls.off[l];
                if (ls.dW[l]) {
                    float* d = ls.dW[l] + (long long)o * K + k;
                    *d = n0 == 0 ? acc[rr] : *d + acc[rr];
                }
                if (ls.db[l] && k == 0) ls.db[l][o] = n0 == 0 ? bs[rr] : ls.db[l][o] + bs[rr];
            }
        }
    }
}
void linear_stack_bwd_w(const float* dy, int ld_dy, const float* x, int ld_x, const LinearStack& ls, int N, int K, cudaStream_t st) {
    const int total = ls.off[ls.n_layers];
    dim3 grid((total + 7) / 8, (K + 255) / 256);
    linear_stack_bwd_w_k<<<grid, 256, 0, st>>>(dy, ld_dy, x, ld_x, ls, N, K);
}
// dx[n, k] = sum over stacked rows r of dy[n, r] * W_l(r)[o(r), k]; grid (N, K / 256), the row loop unrolled by 4
__global__ void __launch_bounds__(256) linear_stack_bwd_x_k(const float* __restrict__ dy, int ld_dy, const __grid_constant__ LinearStack ls,
                                                            float* __restrict__ dx, int ld_dx, int K) {
    const int n = blockIdx.x;
    const int k = blockIdx.y * 256 + threadIdx.x;
    if (k >= K) return;
    const float* dyn = dy + (long long)n * ld_dy;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int l = 0; l < ls.n_layers; ++l) {
        const float* W = ls.W[l] + k;
        const float* g = dyn + ls.off[l];
        const int O = ls.off[l + 1] - ls.off[l];
        int o = 0;
        for (; o + 4 <= O; o += 4) {
            a0 = fmaf(g[o], W[(long long)o * K], a0);
            a1 = fmaf(g[o + 1], W[(long long)(o + 1) * K], a1);
            a2 = fmaf(g[o + 2], W[(long long)(o + 2) * K], a2);
            a3 = fmaf(g[o + 3], W[(long long)(o + 3) * K], a3);
        }
        for (; o < O; ++o) a0 = fmaf(g[o], W[(long long)o * K], a0);
    }
    dx[(long long)n * ld_dx + k] = (a0 + a1) + (a2 + a3);
}
void linear_stack_bwd_x(const float* dy, int ld_dy, const LinearStack& ls, float* dx, int ld_dx, int N, int K, cudaStream_t st) {
    dim3 grid(N, (K + 255) / 256);
    linear_stack_bwd_x_k<<<grid, 256, 0, st>>>(dy, ld_dy, ls, dx, ld_dx, K);
}

__global__ void silu_bwd_mul_k(float* __restrict__ d, const float* __restrict__ x, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float z = x[i];
    const float s = 1.f / (1.f + expf(-z));
    d[i] *= s * (1.f + z * (1.f - s));
}
__global__ void silu_f32_k(const float* __restrict__ x, float* __restrict__ y, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = silu_exact_f(x[i]);
}
void silu_f32(const float* x, float* y, long long n, cudaStream_t st) { silu_f32_k<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n); }
void silu_bwd_mul(float* d, const float* x, long long n, cudaStream_t st) {
    silu_bwd_mul_k<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, x, n);
}

__global__ void conv_out_transpose_weights_k(const float* __restrict__ w, float* __restrict__ w_t, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over [c][o][tap]
    if (i >= C * 27) return;
    const int tap = i % 9, o = (i / 9) % 3, c = i / 27;
    w_t[i] = w[((long long)o * C + c) * 9 + (8 - tap)];
}
void conv_out_transpose_weights(const float* w, float* w_t, int C, cudaStream_t st) {
    conv_out_transpose_weights_k<<<(C * 27 + 255) / 256, 256, 0, st>>>(w, w_t, C);
}
__global__ void conv_out_wgrad_fix_k(const float* __restrict__ t, float* __restrict__ grad, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over grad [o][c][tap]
    if (i >= C * 27) return;
    const int tap = i % 9, c = (i / 9) % C, o = i / (9 * C);
    grad[i] = t[((long long)c * 3 + o) * 9 + (8 - tap)];
}
void conv_out_wgrad_fix(const float* t, float* grad, int C, cudaStream_t st) {
    conv_out_wgrad_fix_k<<<(C * 27 + 255) / 256, 256, 0, st>>>(t, grad, C);
}
// one CTA per (channel, image): out[c] accumulates over images in a second, fixed-order pass (sum_nchw_finish_k)
__global__ void __launch_bounds__(256) sum_nchw_channels_k(const float* __restrict__ x, int N, int C, int HW, float* __restrict__ out) {
    __shared__ float red[8];
    const int c = blockIdx.x, n = blockIdx.y;
    float a = 0.f;
    for (int p = threadIdx.x; p < HW; p += 256) a += x[((long long)n * C + c) * HW + p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int j = 0; j < 8; ++j) s += red[j];
        out[(long long)n * C + c] = s;  // partial [N][C]
    }
}
void sum_nchw_channels(const float* x, int N, int C, int HW, float* ws, float* out, cudaStream_t st) {
    dim3 grid(C, N);
    sum_nchw_channels_k<<<grid, 256, 0, st>>>(x, N, C, HW, ws);
    reduce_rows_f32_k<<<(C + 31) / 32, 256, 0, st>>>(ws, N, C, out);
}

// ================================================================================================ dropout
__device__ __forceinline__ unsigned mix32(unsigned h) {  // murmur3 finaliser
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__global__ void dropout_bf16_k(bf16* __restrict__ x, long long n, float p, unsigned s0, unsigned s1, bf16* __restrict__ mask_out) {
    const float keep_scale = 1.f / (1.f - p);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned lo = (unsigned)i, hi = (unsigned)(i >> 32);
        const unsigned h = mix32(mix32(lo ^ s0) + (hi ^ s1) * 0x9e3779b9u);
        const float u = (float)(h >> 8) * (1.f / 16777216.f);  // uniform in [0, 1)
        const float m = u >= p ? keep_scale : 0.f;
        if (x) x[i] = __float2bfloat16_rn(__bfloat162float(x[i]) * m);
        if (mask_out) mask_out[i] = __float2bfloat16_rn(m);
    }
}
void dropout_bf16(bf16* x, long long n, float p, unsigned long long seed, unsigned stream, bf16* mask_out, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 4736) blocks = 4736;
    const unsigned s0 = (unsigned)seed ^ (stream * 0x9e3779b9u), s1 = (unsigned)(seed >> 32) + stream * 0x85ebca6bu;
    dropout_bf16_k<<<(int)blocks, 256, 0, st>>>(x, n, p, s0, s1, mask_out);
}

// ================================================================================================ ADM / EDM backward helpers
__global__ void __launch_bounds__(256) gn_bwd_film_k(const float2* __restrict__ AB, int N, int C, const float* __restrict__ film, int film_ld,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float* __restrict__ d_film, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float2 red[8][33];
    const int col = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + col;
    float2 a = make_float2(0.f, 0.f);
    if (c < C) {
        const float g = gamma[c], b = beta[c];
        for (int n = rl; n < N; n += 8) {
            const float2 v = AB[(long long)n * C + c];
            const float sc = 1.f + film[(long long)n * film_ld + c];
            d_film[(long long)n * film_ld + c] = fmaf(g, v.y, b * v.x);
            d_film[(long long)n * film_ld + C + c] = v.x;
            a.x = fmaf(sc, v.x, a.x);
            a.y = fmaf(sc, v.y, a.y);
        }
    }
    red[rl][col] = a;
    __syncthreads();
    if (rl == 0 && c < C) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sa += red[j][col].x;
            sb += red[j][col].y;
        }
        if (dbeta) dbeta[c] = sa;
        if (dgamma) dgamma[c] = sb;
    }
}
void gn_bwd_film_params(const float* ws, int N, int HW, int C, const float* film, int film_ld, const float* gamma, const float* beta,
                        float* d_film, float* dgamma, float* dbeta, cudaStream_t st) {
    const float2* AB = reinterpret_cast<const float2*>(ws) + (long long)N * gnb_slabs(HW, C) * C;
    gn_bwd_film_k<<<(C + 31) / 32, 256, 0, st>>>(AB, N, C, film, film_ld, gamma, beta, d_film, dgamma, dbeta);
}

__global__ void __launch_bounds__(256) softmax_rows_k(const float* __restrict__ sc, bf16* __restrict__ P, long long rows, int S) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* x = sc + r * S;
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) m = fmaxf(m, x[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int j = lane; j < S; j += 32) l += __expf(x[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.f / l;
    for (int j = lane; j < S; j += 32) P[r * S + j] = __float2bfloat16_rn(__expf(x[j] - m) * inv);
}
void softmax_rows(const float* scores, bf16* P, long long rows, int S, cudaStream_t st) {
    softmax_rows_k<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(scores, P, rows, S);
}

// attn_small_bwd_k with heads: Ct = heads * d columns per q / k / v block
__global__ void __launch_bounds__(256) attn_small_bwd_heads_k(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                             bf16* __restrict__ dqkv, int S, int d, int Ct, float scale) {
    extern __shared__ uint8_t smb[];
    // rows padded by one 32-bit word: the score loop reads row b of K / V per lane (stride d would be a 32-way bank conflict)
    const int dp = d + 2;
    bf16* sq = reinterpret_cast<bf16*>(smb);  // [S][dp]
    bf16* sk = sq + S * dp;
    bf16* sv = sk + S * dp;
    bf16* sdo = sv + S * dp;
    float* sp = reinterpret_cast<float*>(sdo + S * dp);  // [S][S] probabilities
    float* sds = sp + S * S;                             // [S][S] dP, then dS
    const int n = blockIdx.x, h = blockIdx.y;
    const long long b3 = (long long)n * S * 3 * Ct + (long long)h * d, b1 = (long long)n * S * Ct + (long long)h * d;
    for (int i = threadIdx.x; i < S * d; i += 256) {
        const int t = i / d, c = i % d;
        sq[t * dp + c] = qkv[b3 + (long long)t * 3 * Ct + c];
        sk[t * dp + c] = qkv[b3 + (long long)t * 3 * Ct + Ct + c];
        sv[t * dp + c] = qkv[b3 + (long long)t * 3 * Ct + 2 * Ct + c];
        sdo[t * dp + c] = d_o[b1 + (long long)t * Ct + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * S; i += 256) {
        const int a = i / S, b = i % S;
        float s = 0.f, g = 0.f;
        for (int c = 0; c < d; ++c) {
            s = fmaf(__bfloat162float(sq[a * dp + c]), __bfloat162float(sk[b * dp + c]), s);
            g = fmaf(__bfloat162float(sdo[a * dp + c]), __bfloat162float(sv[b * dp + c]), g);
        }
        sp[i] = s * scale;
        sds[i] = g;  // dP
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int a = threadIdx.x >> 5; a < S; a += 8) {
        float m = -INFINITY;
        for (int b = lane; b < S; b += 32) m = fmaxf(m, sp[a * S + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float l = 0.f;
        for (int b = lane; b < S; b += 32) {
            const float e = __expf(sp[a * S + b] - m);
            sp[a * S + b] = e;
            l += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        const float inv = 1.f / l;
        float t = 0.f;
        for (int b = lane; b < S; b += 32) {
            const float p = sp[a * S + b] * inv;
            sp[a * S + b] = p;
            t = fmaf(p, sds[a * S + b], t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        for (int b = lane; b < S; b += 32) sds[a * S + b] = sp[a * S + b] * (sds[a * S + b] - t) * scale;  // dS
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * d; i += 256) {
        const int t = i / d, c = i % d;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < S; ++j) {
            dq = fmaf(sds[t * S + j], __bfloat162float(sk[j * dp + c]), dq);
            dk = fmaf(sds[j * S + t], __bfloat162float(sq[j * dp + c]), dk);
            dv = fmaf(sp[j * S + t], __bfloat162float(sdo[j * dp + c]), dv);
        }
        dqkv[b3 + (long long)t * 3 * Ct + c] = __float2bfloat16_rn(dq);
        dqkv[b3 + (long long)t * 3 * Ct + Ct + c] = __float2bfloat16_rn(dk);
        dqkv[b3 + (long long)t * 3 * Ct + 2 * Ct + c] = __float2bfloat16_rn(dv);
    }
}
void attn_small_bwd_heads(const bf16* qkv, const bf16* d_o, bf16* dqkv, int N, int heads, int S, int d, float scale, cudaStream_t st) {
    const size_t smem = (size_t)4 * S * (d + 2) * sizeof(bf16) + (size_t)2 * S * S * sizeof(float);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(attn_small_bwd_heads_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured.set();
    }
    attn_small_bwd_heads_k<<<dim3(N, heads), 256, smem, st>>>(qkv, d_o, dqkv, S, d, heads * d, scale);
}

__global__ void embedding_grad_k(const float* __restrict__ d_emb, const long long* __restrict__ idx, float* __restrict__ grad, int N, int D,
                                 int num_classes) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D) return;
    for (int n = 0; n < N; ++n) {
        const long long y = idx[n];
        if (y >= 0 && y < num_classes) grad[y * D + k] += d_emb[(long long)n * D + k];
    }
}
void embedding_grad(const float* d_emb, const long long* idx, float* grad, int N, int D, int num_classes, cudaStream_t st) {
    cudaMemsetAsync(grad, 0, (size_t)num_classes * D * sizeof(float), st);
    embedding_grad_k<<<(D + 127) / 128, 128, 0, st>>>(d_emb, idx, grad, N, D, num_classes);
}

__global__ void scale_bf16_k(bf16* __restrict__ x, long long n8, float s) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float f[8];
        unpack8(ld8(x + i * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= s;
        st8(x + i * 8, pack8(f));
    }
}
void scale_bf16(bf16* x, long long n, float s, cudaStream_t st) {
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 2368) blocks = 2368;
    if (blocks < 1) blocks = 1;
    scale_bf16_k<<<(int)blocks, 256, 0, st>>>(x, n / 8, s);
}

}  // namespace dxmi
