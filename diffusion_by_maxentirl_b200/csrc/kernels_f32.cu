// fp32 mode kernels (see kernels_f32.cuh): plain CUDA-core arithmetic, exact expf / division, fixed reduction orders.
#include "kernels_f32.cuh"

#include <cstdint>

namespace dxmi {

__device__ __forceinline__ float silu_exact(float v) { return v / (1.f + expf(-v)); }

// ================================================================================================ convolution
// Implicit GEMM on CUDA cores: CTA tile 64 output pixels x 64 output channels, K = taps x input channels in slices of 16;
// each of the 256 threads owns a 4 x 4 micro tile.
static constexpr int CV_BM = 64, CV_BN = 64, CV_BK = 16;

__global__ void __launch_bounds__(256) conv_f32_k(const ConvF32 c, int Ho, int Wo, long long M) {
    __shared__ float As[CV_BK][CV_BM + 4];
    __shared__ float Bs[CV_BK][CV_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int Cin = c.C1 + c.C2;
    const int kk2 = c.k * c.k;
    const int pad = c.stride == 1 ? (c.k - 1) / 2 : 0;
    const long long m0 = (long long)blockIdx.x * CV_BM;
    const int co0 = blockIdx.y * CV_BN;

    // loader roles: this thread fetches 4 consecutive K entries of one pixel (A) and of one output channel (B)
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const long long lm = m0 + lrow;
    const bool lm_ok = lm < M;
    int ln = 0, loh = 0, low = 0;
    if (lm_ok) {
        ln = (int)(lm / ((long long)Ho * Wo));
        const int rem = (int)(lm - (long long)ln * Ho * Wo);
        loh = rem / Wo;
        low = rem - loh * Wo;
    }
    const int lco = co0 + lrow;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < kk2; ++tap) {
        const int r = tap / c.k, q = tap - r * c.k;
        const int ih = loh * c.stride + r - pad, iw = low * c.stride + q - pad;
        const bool pix_ok = lm_ok && ih >= 0 && ih < c.H && iw >= 0 && iw < c.W;
        const long long pix = ((long long)ln * c.H + ih) * c.W + iw;
        for (int c0 = 0; c0 < Cin; c0 += CV_BK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = c0 + lk + j;
                float a = 0.f, b = 0.f;
                if (ch < Cin) {
                    if (pix_ok) a = ch < c.C1 ? c.x1[pix * c.C1 + ch] : c.x2[pix * c.C2 + (ch - c.C1)];
                    if (lco < c.Cout) b = c.w[((long long)lco * Cin + ch) * kk2 + tap];
                }
                As[lk + j][lrow] = a;
                Bs[lk + j][lrow] = b;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < CV_BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int n = (int)(m / ((long long)Ho * Wo));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co >= c.Cout) continue;
            float v = acc[i][j];
            if (c.bias) v += c.bias[co];
            if (c.rowvec) v += c.rowvec[(long long)n * c.ldrv + co];
            if (c.residual) v += c.residual[m * c.Cout + co];
            if (c.act == 1) v = v > 0.f ? v : 0.2f * v;
            c.out[m * c.Cout + co] = v;
        }
    }
}

void conv_f32(const ConvF32& c, cudaStream_t st) {
    const int Ho = c.stride == 1 ? c.H : c.H / 2, Wo = c.stride == 1 ? c.W : c.W / 2;
    const long long M = (long long)c.N * Ho * Wo;
    dim3 grid((unsigned)((M + CV_BM - 1) / CV_BM), (unsigned)((c.Cout + CV_BN - 1) / CV_BN));
    conv_f32_k<<<grid, 256, 0, st>>>(c, Ho, Wo, M);
}

// ================================================================================================ GroupNorm
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // red may still be read from a previous call
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(256) group_norm_f32_k(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                                       int C2, int HW, int groups, float eps, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int silu, float* __restrict__ out) {
    __shared__ float red[8];
    const int C = C1 + C2, cpg = C / groups;
    const int g = blockIdx.x, n = blockIdx.y;
    const int total = HW * cpg;
    auto load = [&](int e) {
        const int p = e / cpg, ch = g * cpg + (e - p * cpg);
        const long long pix = (long long)n * HW + p;
        return ch < C1 ? x1[pix * C1 + ch] : x2[pix * C2 + (ch - C1)];
    };
    float s = 0.f;
    for (int e = threadIdx.x; e < total; e += 256) s += load(e);
    const float mean = block_sum_256(s, red) / (float)total;
    float q = 0.f;
    for (int e = threadIdx.x; e < total; e += 256) {
        const float d = load(e) - mean;
        q = fmaf(d, d, q);
    }
    const float var = block_sum_256(q, red) / (float)total;
    const float rstd = 1.f / sqrtf(var + eps);
    for (int e = threadIdx.x; e < total; e += 256) {
        const int p = e / cpg, ch = g * cpg + (e - p * cpg);
        float v = (load(e) - mean) * rstd * gamma[ch] + beta[ch];
        if (silu) v = silu_exact(v);
        out[((long long)n * HW + p) * C + ch] = v;
    }
}

void group_norm_f32(const float* x1, int C1, const float* x2, int C2, int N, int HW, int groups, float eps,
                    const float* gamma, const float* beta, int silu, float* out, cudaStream_t st) {
    dim3 grid(groups, N);
    group_norm_f32_k<<<grid, 256, 0, st>>>(x1, C1, x2, C2, HW, groups, eps, gamma, beta, silu, out);
}

// ================================================================================================ attention
static constexpr int AT_QB = 8;  // queries per CTA (one warp per query in the softmax)

__global__ void __launch_bounds__(256) attention_f32_k(const float* __restrict__ q, const float* __restrict__ k,
                                                      const float* __restrict__ v, float* __restrict__ out, int HW, int C,
                                                      float scale) {
    extern __shared__ float sm[];
    float* sq = sm;               // [AT_QB][C]
    float* sp = sm + AT_QB * C;   // [AT_QB][HW]
    const int n = blockIdx.y, q0 = blockIdx.x * AT_QB;
    const long long base = (long long)n * HW;
    for (int i = threadIdx.x; i < AT_QB * C; i += 256) sq[i] = q[(base + q0 + i / C) * C + (i % C)];
    __syncthreads();
    for (int j = threadIdx.x; j < HW; j += 256) {
        float d[AT_QB];
#pragma unroll
        for (int i = 0; i < AT_QB; ++i) d[i] = 0.f;
        const float* kr = k + (base + j) * C;
        for (int cc = 0; cc < C; ++cc) {
            const float kv = kr[cc];
#pragma unroll
            for (int i = 0; i < AT_QB; ++i) d[i] = fmaf(sq[i * C + cc], kv, d[i]);
        }
#pragma unroll
        for (int i = 0; i < AT_QB; ++i) sp[i * HW + j] = d[i] * scale;
    }
    __syncthreads();
    {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float* row = sp + w * HW;
        float mx = -INFINITY;
        for (int j = lane; j < HW; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int j = lane; j < HW; j += 32) {
            const float e = expf(row[j] - mx);
            row[j] = e;
            s += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        for (int j = lane; j < HW; j += 32) row[j] = row[j] / s;
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < C; cc += 256) {
        float a[AT_QB];
#pragma unroll
        for (int i = 0; i < AT_QB; ++i) a[i] = 0.f;
        for (int j = 0; j < HW; ++j) {
            const float vv = v[(base + j) * C + cc];
#pragma unroll
            for (int i = 0; i < AT_QB; ++i) a[i] = fmaf(sp[i * HW + j], vv, a[i]);
        }
#pragma unroll
        for (int i = 0; i < AT_QB; ++i) out[(base + q0 + i) * C + cc] = a[i];
    }
}

void attention_f32(const float* q, const float* k, const float* v, float* out, int N, int HW, int C, float scale,
                   cudaStream_t st) {
    dim3 grid(HW / AT_QB, N);
    const size_t smem = (size_t)AT_QB * (C + HW) * sizeof(float);
    attention_f32_k<<<grid, 256, smem, st>>>(q, k, v, out, HW, C, scale);
}

// ================================================================================================ small helpers
__global__ void linear_exact_f32_k(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                   const float* __restrict__ b, float* __restrict__ y, int ldy, int N, int K, int O,
                                   int act_in) {
    const int lane = threadIdx.x & 31;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)N * O) return;
    const int n = (int)(wid / O), o = (int)(wid % O);
    float a = 0.f;
    for (int kk = lane; kk < K; kk += 32) {
        float xv = x[(long long)n * ldx + kk];
        if (act_in == 2) xv = silu_exact(xv);
        a = fmaf(xv, W[(long long)o * K + kk], a);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) y[(long long)n * ldy + o] = a + (b ? b[o] : 0.f);
}
void linear_exact_f32(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int N, int K, int O,
                      int act_in, cudaStream_t st) {
    const long long threads = (long long)N * O * 32;
    linear_exact_f32_k<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, ldx, W, b, y, ldy, N, K, O, act_in);
}

__global__ void upsample2x_f32_k(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C) {
    const long long total = (long long)N * 4 * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C);
        long long r = i / C;
        const int ow = (int)(r % (2 * W));
        r /= 2 * W;
        const int oh = (int)(r % (2 * H));
        const int n = (int)(r / (2 * H));
        out[i] = x[(((long long)n * H + (oh >> 1)) * W + (ow >> 1)) * C + ch];
    }
}
void upsample2x_f32(const float* x, float* out, int N, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)N * 4 * H * W * C;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    upsample2x_f32_k<<<(unsigned)blocks, 256, 0, st>>>(x, out, N, H, W, C);
}

__global__ void avgpool2_f32_k(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C, int act) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C);
        long long r = i / C;
        const int ow = (int)(r % Wo);
        r /= Wo;
        const int oh = (int)(r % Ho);
        const int n = (int)(r / Ho);
        const float* p = x + (((long long)n * H + 2 * oh) * W + 2 * ow) * C + ch;
        float v = ((p[0] + p[C]) + (p[(long long)W * C] + p[(long long)W * C + C])) * 0.25f;
        if (act == 1) v = v > 0.f ? v : 0.2f * v;
        out[i] = v;
    }
}
void avgpool2_f32(const float* x, float* out, int N, int H, int W, int C, int act, cudaStream_t st) {
    const long long total = (long long)N * (H / 2) * (W / 2) * C;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    avgpool2_f32_k<<<(unsigned)blocks, 256, 0, st>>>(x, out, N, H, W, C, act);
}

__global__ void nchw_to_nhwc_f32_k(const float* __restrict__ x, float* __restrict__ out, int C, int HW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C);
        const long long r = i / C;
        const int p = (int)(r % HW);
        const long long n = r / HW;
        out[i] = x[(n * C + ch) * HW + p];
    }
}
void nchw_to_nhwc_f32(const float* x, float* out, int N, int C, int HW, cudaStream_t st) {
    const long long total = (long long)N * C * HW;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    nchw_to_nhwc_f32_k<<<(unsigned)blocks, 256, 0, st>>>(x, out, C, HW, total);
}
__global__ void nhwc_to_nchw_f32_k(const float* __restrict__ x, float* __restrict__ out, int C, int HW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const long long r = i / HW;
        const int ch = (int)(r % C);
        const long long n = r / C;
        out[i] = x[(n * HW + p) * C + ch];
    }
}
void nhwc_to_nchw_f32(const float* x, float* out, int N, int C, int HW, cudaStream_t st) {
    const long long total = (long long)N * C * HW;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    nhwc_to_nchw_f32_k<<<(unsigned)blocks, 256, 0, st>>>(x, out, C, HW, total);
}

// one CTA per image: channel c = thread; relu, sum over HW in pixel order, dot with the Linear(C,1) weight
__global__ void __launch_bounds__(256) value_head_f32_k(const float* __restrict__ h, int HW, int C,
                                                       const float* __restrict__ lin_w, const float* __restrict__ lin_b,
                                                       const float* __restrict__ scale_w, const float* __restrict__ scale_b,
                                                       float* __restrict__ out) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float part = 0.f;
    for (int ch = threadIdx.x; ch < C; ch += 256) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += fmaxf(h[((long long)n * HW + p) * C + ch], 0.f);
        part = fmaf(s, lin_w[ch], part);
    }
    const float tot = block_sum_256(part, red);
    if (threadIdx.x == 0) {
        float v = tot + lin_b[0];
        if (scale_w) v = v * scale_w[0] + scale_b[0];
        out[n] = v;
    }
}
void value_head_f32(const float* h, int N, int HW, int C, const float* lin_w, const float* lin_b, const float* scale_w,
                    const float* scale_b, float* out, cudaStream_t st) {
    value_head_f32_k<<<N, 256, 0, st>>>(h, HW, C, lin_w, lin_b, scale_w, scale_b, out);
}

}  // namespace dxmi
