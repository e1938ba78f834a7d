// HBM-bound / small kernels of the DxMI sampler path. See kernels.cuh for contracts.
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <cuda_fp16.h>
#include <math_constants.h>

namespace dxmi {

namespace {

// silu(v) = v * sigmoid(v) = h + h * tanh(h) with h = v / 2: ONE special-function op (tanh.approx, rel. error ~2^-11, well
// below the bf16 output rounding) instead of exp + reciprocal - the GroupNorm apply kernels were SFU bound (ncu: XU pipe
// 62 % vs DRAM 43 %).
__device__ __forceinline__ float silu_f(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ float act_f(float v, int act) {
    if (act == 1) return v > 0.f ? v : 0.2f * v;
    if (act == 2) return silu_f(v);
    return v;
}

struct alignas(16) bf16x8 {
    __nv_bfloat162 v[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(p.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
    bf16x8 p;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return p;
}
// 16-byte vector access. (Copying the struct member-wise - what `*reinterpret_cast<const bf16x8*>(p)` compiles to, because
// __nv_bfloat162 has user-provided copy operations - became FOUR 4-byte LDG / STG per thread: ncu round 2 showed 8.9 of 32
// bytes used per sector in the GroupNorm kernels.  Go through uint4 so it is one LDG.128 / STG.128.)
__device__ __forceinline__ bf16x8 ld8(const bf16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    bf16x8 r;
    *reinterpret_cast<uint4*>(&r) = u;
    return r;
}
__device__ __forceinline__ void st8(bf16* p, const bf16x8& v) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v); }

// launch with programmatic stream serialization (PDL, ptx.cuh): the kernel MUST call ptx::pdl_wait() before its first global access
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = (l < (blockDim.x >> 5)) ? sh[l] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

}  // namespace

// ============================================================================================ weight packing

template <typename T>
__global__ void pack_conv_weight_k(const T* __restrict__ w, int Cout, int Cin, int taps, int c_off, int c_cnt,
                                   bf16* __restrict__ dst, long long ldk, long long k_off) {
    const long long total = (long long)Cout * taps * c_cnt;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c_cnt);
        const long long r = i / c_cnt;
        const int tap = (int)(r % taps);
        const int o = (int)(r / taps);
        const float v = (float)w[((long long)o * Cin + c_off + cc) * taps + tap];
        dst[o * ldk + k_off + (long long)tap * c_cnt + cc] = __float2bfloat16_rn(v);
    }
}

void pack_conv_weight(const void* w, int w_is_half, int Cout, int Cin, int kh, int kw, int c_off, int c_cnt, bf16* dst,
                      long long ldk, long long k_off, cudaStream_t st) {
    const long long total = (long long)Cout * kh * kw * c_cnt;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    if (w_is_half)
        pack_conv_weight_k<__half><<<blocks, 256, 0, st>>>((const __half*)w, Cout, Cin, kh * kw, c_off, c_cnt, dst, ldk, k_off);
    else
        pack_conv_weight_k<float><<<blocks, 256, 0, st>>>((const float*)w, Cout, Cin, kh * kw, c_off, c_cnt, dst, ldk, k_off);
}

// nearest-2x upsample + 3x3 conv = four 2x2 phase convolutions of the low-resolution input: output row Y = 2y + py reads upsampled rows
// Y - 1 .. Y + 1 = low-resolution rows {y - 1, y, y} (py = 0) or {y, y, y + 1} (py = 1), so tap dy of phase py sums kernel rows
// ky in {0} | {1, 2} (py = 0) or {0, 1} | {2} (py = 1); columns likewise.  Sums in fp32, one bf16 rounding.
template <typename T>
__global__ void pack_conv_weight_up2_k(const T* __restrict__ w, int Cout, int Cin, bf16* __restrict__ dst) {
    const long long total = 4LL * Cout * 4 * Cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        long long r = i / Cin;
        const int tap = (int)(r & 3);
        r >>= 2;
        const int o = (int)(r % Cout);
        const int phase = (int)(r / Cout);
        const int py = phase >> 1, px = phase & 1, dy = tap >> 1, dx = tap & 1;
        const int ky0 = py == 0 ? (dy == 0 ? 0 : 1) : (dy == 0 ? 0 : 2), ky1 = py == 0 ? (dy == 0 ? 0 : 2) : (dy == 0 ? 1 : 2);
        const int kx0 = px == 0 ? (dx == 0 ? 0 : 1) : (dx == 0 ? 0 : 2), kx1 = px == 0 ? (dx == 0 ? 0 : 2) : (dx == 0 ? 1 : 2);
        const T* wp = w + ((long long)o * Cin + c) * 9;
        float v = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) v += (float)wp[ky * 3 + kx];
        dst[i] = __float2bfloat16_rn(v);
    }
}
void pack_conv_weight_up2(const void* w, int w_is_half, int Cout, int Cin, bf16* dst, cudaStream_t st) {
    const long long total = 16LL * Cout * Cin;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    if (w_is_half)
        pack_conv_weight_up2_k<__half><<<blocks, 256, 0, st>>>((const __half*)w, Cout, Cin, dst);
    else
        pack_conv_weight_up2_k<float><<<blocks, 256, 0, st>>>((const float*)w, Cout, Cin, dst);
}

// transposed + tap-flipped packing for the data gradient: dX = conv(dY, W') with W'[ci][tap'][co] = W[co][ci][taps-1-tap']
template <typename T>
__global__ void pack_conv_weight_dgrad_k(const T* __restrict__ w, int Cout, int Cin, int taps, bf16* __restrict__ dst,
                                         long long ldk, long long k_off) {
    const long long total = (long long)Cin * taps * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long r = i / Cout;
        const int tap = (int)(r % taps);
        const int ci = (int)(r / taps);
        const float v = (float)w[((long long)co * Cin + ci) * taps + (taps - 1 - tap)];
        dst[ci * ldk + k_off + (long long)tap * Cout + co] = __float2bfloat16_rn(v);
    }
}
void pack_conv_weight_dgrad(const void* w, int w_is_half, int Cout, int Cin, int taps, bf16* dst, long long ldk, long long k_off,
                            cudaStream_t st) {
    const long long total = (long long)Cout * taps * Cin;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    if (w_is_half)
        pack_conv_weight_dgrad_k<__half><<<blocks, 256, 0, st>>>((const __half*)w, Cout, Cin, taps, dst, ldk, k_off);
    else
        pack_conv_weight_dgrad_k<float><<<blocks, 256, 0, st>>>((const float*)w, Cout, Cin, taps, dst, ldk, k_off);
}

template <typename T>
__global__ void cast_to_f32_k(const T* __restrict__ s, float* __restrict__ d, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        d[i] = (float)s[i];
}
void cast_to_f32(const void* src, int src_is_half, float* dst, long long n, cudaStream_t st) {
    const int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
    if (src_is_half)
        cast_to_f32_k<__half><<<blocks, 256, 0, st>>>((const __half*)src, dst, n);
    else
        cast_to_f32_k<float><<<blocks, 256, 0, st>>>((const float*)src, dst, n);
}

// ============================================================================================ first / last conv

// Persistent CTAs loop over tiles of 128 consecutive output pixels of one image (th rows x tw columns).  The weights
// ([tap*Cin + ci][Cout], fp32) are staged in shared memory once per CTA, the zero-padded fp32 input patch once per tile.
// A thread owns 8 consecutive pixels x 8 output channels: the 16 lanes that share a pixel group write one full
// Cout-wide NHWC line per store (coalesced), every input value read from smem feeds 8 FMAs and every weight vector 32.
// Optionally emits the GroupNorm partial statistics of its bf16 outputs in the layout of the GEMM epilogues
// ([N][HW/128][Cout][2], one partial per tile) so that the consumer GroupNorm needs no statistics pass.
template <int CIN>
__global__ void __launch_bounds__(512) conv3x3_first_k(const float* __restrict__ x, const float* __restrict__ in_scale,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      bf16* __restrict__ out, float* __restrict__ stats, int H, int W, int Cout,
                                                      int act, int th, int tw, int ntiles) {
    extern __shared__ float sm[];
    constexpr int K = 9 * CIN;
    const int pw = tw + 2, ph = th + 2;
    float* sw = sm;                        // [K][Cout]
    float* sb = sw + K * Cout;             // [Cout]
    float* sx = sb + Cout;                 // [CIN][ph][pw]
    const int patch = (CIN * ph * pw + 3) & ~3;
    float2* red = reinterpret_cast<float2*>(sx + 2 * patch);  // [16][Cout]  (sx is double buffered)
    const int tiles_w = W / tw, tiles_per_img = tiles_w * (H / th);
    const long long HW = (long long)H * W;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int o = i % Cout, kk = i / Cout;  // kk = tap*CIN + ci
        const int tap = kk / CIN, ci = kk % CIN;
        sw[i] = w[((long long)o * CIN + ci) * 9 + tap];
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = b ? b[i] : 0.f;
    const int CG = Cout >> 3;
    const int cg = threadIdx.x % CG, pg = threadIdx.x / CG;  // channels cg*8..+7, tile pixels pg*8..+7
    const int pr = (pg * 8) / tw, pc0 = (pg * 8) % tw;

    // The zero-padded input patch of tile i + 1 is fetched into registers BEFORE tile i is computed and parked in the other half of
    // sx afterwards: the first version paid one exposed global-load latency per tile (72 us per launch at 32x32, B = 256, against
    // 25 us of FMA work).
    constexpr int NPRE = 6;  // >= ceil(CIN * ph * pw / blockDim): 3 (32-wide maps), 4 (64), 5 (256)
    float nx[NPRE];
    auto fetch = [&](int tile) {
        const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
        const int h0 = (trem / tiles_w) * th, w0 = (trem % tiles_w) * tw;
        const float sc = in_scale ? in_scale[n] : 1.f;
#pragma unroll
        for (int u = 0; u < NPRE; ++u) {
            const int i = threadIdx.x + u * blockDim.x;
            float v = 0.f;
            if (i < CIN * ph * pw) {
                const int ci = i / (ph * pw), r = (i / pw) % ph, c = i % pw;
                const int hh = h0 + r - 1, ww = w0 + c - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((long long)n * CIN + ci) * HW + (long long)hh * W + ww] * sc;
            }
            nx[u] = v;
        }
    };
    auto park = [&](float* dst) {
#pragma unroll
        for (int u = 0; u < NPRE; ++u) {
            const int i = threadIdx.x + u * blockDim.x;
            if (i < CIN * ph * pw) dst[i] = nx[u];
        }
    };
    const bool pre_ok = CIN * ph * pw <= NPRE * (int)blockDim.x;  // (narrow nets: few threads per CTA - stage synchronously)
    if (pre_ok && blockIdx.x < ntiles) {
        fetch(blockIdx.x);
        park(sx);
    }
    int buf = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
        const int h0 = (trem / tiles_w) * th, w0 = (trem % tiles_w) * tw;
        __syncthreads();  // this tile's patch is parked; the previous tile's readers of the other buffer / red are done (and sw / sb are staged)
        const float* sxc = sx + buf * patch;
        const bool more = pre_ok && tile + gridDim.x < ntiles;
        if (more) fetch(tile + gridDim.x);
        if (!pre_ok) {
            const float sc = in_scale ? in_scale[n] : 1.f;
            for (int i = threadIdx.x; i < CIN * ph * pw; i += blockDim.x) {
                const int ci = i / (ph * pw), r = (i / pw) % ph, c = i % pw;
                const int hh = h0 + r - 1, ww = w0 + c - 1;
                sx[buf * patch + i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[((long long)n * CIN + ci) * HW + (long long)hh * W + ww] * sc : 0.f;
            }
            __syncthreads();
        }
        float acc[8][8];
        {
            const float4 b0 = *reinterpret_cast<const float4*>(sb + cg * 8), b1 = *reinterpret_cast<const float4*>(sb + cg * 8 + 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
                acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
            }
        }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) {
                float xv[10];
                const float* xr = sxc + (ci * ph + pr + dr) * pw + pc0;
#pragma unroll
                for (int j = 0; j < 10; ++j) xv[j] = xr[j];
#pragma unroll
                for (int dc = 0; dc < 3; ++dc) {
                    const float* wr = sw + ((dr * 3 + dc) * CIN + ci) * Cout + cg * 8;
                    const float4 w0v = *reinterpret_cast<const float4*>(wr), w1v = *reinterpret_cast<const float4*>(wr + 4);
                    const float wv[8] = {w0v.x, w0v.y, w0v.z, w0v.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(xv[j + dc], wv[c], acc[j][c]);
                }
            }
        }
        if (more) park(sx + (buf ^ 1) * patch);  // (its last readers finished before this iteration's barrier)
        float s1[8], s2[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) s1[c] = s2[c] = 0.f;
        const long long p0 = (long long)n * HW + (long long)(h0 + pr) * W + w0 + pc0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float f[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) f[c] = act_f(acc[j][c], act);
            const bf16x8 v = pack8(f);
            st8(out + (p0 + j) * Cout + cg * 8, v);
            if (stats) {
                float r[8];
                unpack8(v, r);  // statistics of the values the consumer reads (bf16-rounded)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    s1[c] += r[c];
                    s2[c] = fmaf(r[c], r[c], s2[c]);
                }
            }
        }
        if (stats) {
#pragma unroll
            for (int c = 0; c < 8; ++c) red[pg * Cout + cg * 8 + c] = make_float2(s1[c], s2[c]);
            __syncthreads();
            for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
                float2 a = make_float2(0.f, 0.f);
#pragma unroll
                for (int g = 0; g < 16; ++g) {  // fixed order: deterministic, batch invariant
                    const float2 v = red[g * Cout + c];
                    a.x += v.x;
                    a.y += v.y;
                }
                *reinterpret_cast<float2*>(stats + ((long long)tile * Cout + c) * 2) = a;
            }
        }
    }
}

void conv3x3_first(const float* x, const float* in_scale, const float* w, const float* b, bf16* out, float* stats, int N,
                   int Cin, int H, int W, int Cout, int act, cudaStream_t st) {
    // requirements (checked by the plan builders): Cin == 3, Cout % 32 == 0, Cout <= 256, H*W % 128 == 0, W power of two >= 8
    const int tw = W < 128 ? W : 128;
    const int th = 128 / tw;
    const int threads = 16 * (Cout / 8);
    const int patch = (Cin * (th + 2) * (tw + 2) + 3) & ~3;
    const size_t smem = (size_t)(9 * Cin * Cout + Cout + 2 * patch) * sizeof(float) + (size_t)16 * Cout * sizeof(float2);
    static DevFlags configured;
    static int num_sms = 148;
    if (!configured.test()) {
        cudaFuncSetAttribute(conv3x3_first_k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.set();
    }
    const int ntiles = N * (H / th) * (W / tw);
    const int per_sm = threads <= 256 ? 2 : 1;
    int blocks = num_sms * per_sm;
    if (blocks > ntiles) blocks = ntiles;
    conv3x3_first_k<3><<<blocks, threads, smem, st>>>(x, in_scale, w, b, out, stats, H, W, Cout, act, th, tw, ntiles);
}

// One warp per output pixel; lanes stride over (tap, 8-channel vector) work items. Weights in smem [o][tap][C].
__global__ void conv3x3_last_k(const bf16* __restrict__ hin, const float* __restrict__ w, const float* __restrict__ b,
                               float* __restrict__ out, int N, int C, int H, int W, int Cout) {
    extern __shared__ float sw[];  // [Cout][9][C]
    for (int i = threadIdx.x; i < Cout * 9 * C; i += blockDim.x) {
        const int c = i % C;
        const int tap = (i / C) % 9;
        const int o = i / (9 * C);
        sw[i] = w[((long long)o * C + c) * 9 + tap];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long HW = (long long)H * W;
    const long long total = (long long)N * HW;
    const int CV = C / 8;
    for (long long p = blockIdx.x * (long long)warps_per_block + (threadIdx.x >> 5); p < total;
         p += (long long)gridDim.x * warps_per_block) {
        const int n = (int)(p / HW);
        const int hw = (int)(p % HW);
        const int h = hw / W, wx = hw % W;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int item = lane; item < 9 * CV; item += 32) {
            const int tap = item / CV, cv = item % CV;
            const int hh = h + tap / 3 - 1, ww = wx + tap % 3 - 1;
            if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            float f[8];
            unpack8(ld8(hin + ((long long)n * HW + (long long)hh * W + ww) * C + cv * 8), f);
            for (int o = 0; o < Cout; ++o) {
                const float* wr = sw + (o * 9 + tap) * C + cv * 8;
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fmaf(f[j], wr[j], s);
                acc[o] += s;
            }
        }
        for (int o = 0; o < Cout; ++o) {
            const float s = warp_sum(acc[o]);
            if (lane == 0) out[((long long)n * Cout + o) * HW + hw] = s + (b ? b[o] : 0.f);
        }
    }
}

void conv3x3_last(const bf16* h, const float* w, const float* b, float* out, int N, int C, int H, int W, int Cout,
                  cudaStream_t st) {
    const long long total = (long long)N * H * W;
    long long blocks = (total + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = (size_t)Cout * 9 * C * sizeof(float);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(conv3x3_last_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        configured.set();
    }
    conv3x3_last_k<<<(int)blocks, 256, smem, st>>>(h, w, b, out, N, C, H, W, Cout);
}

// ============================================================================================ GroupNorm

int gn_num_slabs(int N, int HW) {
    // a function of the image size only: the partial-sum order (hence every bit of the result) must not depend on the
    // batch size - bf16 rounding downstream amplifies even 1e-7 differences in the statistics to ~1e-3 at the output
    (void)N;
    int slabs = 1;
    while (slabs < 16 && HW / (slabs * 2) >= 64) slabs *= 2;
    return slabs;
}

static constexpr int GN_THREADS = 256;

__device__ __forceinline__ const bf16* gn_src(const bf16* x1, int C1, int ld1, const bf16* x2, int ld2, long long pix,
                                              int c) {
    return (c < C1) ? x1 + pix * ld1 + c : x2 + pix * ld2 + (c - C1);
}

__global__ void __launch_bounds__(GN_THREADS) gn_stats_k(const bf16* __restrict__ x1, int C1, int ld1,
                                                        const bf16* __restrict__ x2, int C2, int ld2, int HW, int groups,
                                                        float* __restrict__ partial, int slabs) {
    __shared__ float s_sum[GN_THREADS * 8];
    __shared__ float s_sq[GN_THREADS * 8];
    const int C = C1 + C2;
    const int CV = C / 8;
    const int PL = GN_THREADS / CV;
    const int n = blockIdx.x, slab = blockIdx.y;
    const int pix_per_slab = HW / slabs;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    float sum[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
    if (pl < PL) {
        const long long base = (long long)n * HW + (long long)slab * pix_per_slab;
        for (int p = pl; p < pix_per_slab; p += PL) {
            float f[8];
            unpack8(ld8(gn_src(x1, C1, ld1, x2, ld2, base + p, cv * 8)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sum[j] += f[j];
                sq[j] = fmaf(f[j], f[j], sq[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s_sum[pl * C + cv * 8 + j] = sum[j];
            s_sq[pl * C + cv * 8 + j] = sq[j];
        }
    }
    __syncthreads();
    // per-group reduction in a fixed order (deterministic): 2 * groups threads
    const int cpg = C / groups;
    if (threadIdx.x < 2 * groups) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        const float* src = which ? s_sq : s_sum;
        float a = 0.f;
        for (int q = 0; q < PL; ++q)
            for (int c = 0; c < cpg; ++c) a += src[q * C + g * cpg + c];
        partial[((long long)n * slabs + slab) * (2 * groups) + g * 2 + which] = a;
    }
}

void gn_stats(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups,
              float* partial, int slabs, cudaStream_t st) {
    dim3 grid(N, slabs);
    gn_stats_k<<<grid, GN_THREADS, 0, st>>>(x1, C1, ld1, x2, C2, ld2, HW, groups, partial, slabs);
}

__global__ void __launch_bounds__(GN_THREADS) gn_apply_k(const bf16* __restrict__ x1, int C1, int ld1,
                                                        const bf16* __restrict__ x2, int C2, int ld2, int HW, int groups,
                                                        float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ film,
                                                        int film_ld, int silu, const float* __restrict__ partial,
                                                        int slabs, bf16* __restrict__ out) {
    __shared__ float s_mean[64], s_rstd[64];
    const int C = C1 + C2;
    const int CV = C / 8;
    const int PL = GN_THREADS / CV;
    const int n = blockIdx.x, slab = blockIdx.y;
    const int pix_per_slab = HW / slabs;
    const int cpg = C / groups;
    if (threadIdx.x < groups) {
        float s = 0.f, q = 0.f;
        for (int i = 0; i < slabs; ++i) {
            s += partial[((long long)n * slabs + i) * (2 * groups) + threadIdx.x * 2];
            q += partial[((long long)n * slabs + i) * (2 * groups) + threadIdx.x * 2 + 1];
        }
        const float cnt = (float)cpg * (float)HW;
        const float mean = s / cnt;
        const float var = fmaxf(q / cnt - mean * mean, 0.f);
        s_mean[threadIdx.x] = mean;
        s_rstd[threadIdx.x] = rsqrtf(var + eps);
    }
    __syncthreads();
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    if (pl >= PL) return;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j;
        const int g = c / cpg;
        float aa = s_rstd[g] * gamma[c];
        float bb = beta[c] - s_mean[g] * aa;
        if (film) {
            const float sc = 1.f + film[(long long)n * film_ld + c];
            const float sh = film[(long long)n * film_ld + C + c];
            aa *= sc;
            bb = bb * sc + sh;
        }
        a[j] = aa;
        b[j] = bb;
    }
    const long long base = (long long)n * HW + (long long)slab * pix_per_slab;
    for (int p = pl; p < pix_per_slab; p += PL) {
        float f[8];
        unpack8(ld8(gn_src(x1, C1, ld1, x2, ld2, base + p, cv * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = fmaf(f[j], a[j], b[j]);
            f[j] = silu ? silu_f(v) : v;
        }
        st8(out + (base + p) * C + cv * 8, pack8(f));
    }
}

void gn_apply(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups, float eps,
              const float* gamma, const float* beta, const float* film, int film_ld, int silu, const float* partial,
              int slabs, bf16* out, cudaStream_t st) {
    dim3 grid(N, slabs);
    gn_apply_k<<<grid, GN_THREADS, 0, st>>>(x1, C1, ld1, x2, C2, ld2, HW, groups, eps, gamma, beta, film, film_ld, silu,
                                             partial, slabs, out);
}

// GroupNorm apply with the statistics taken from the producer GEMMs' fused partials (gemm_tc2.cu):
// st1 / st2 = [N][P1 | P2][C1 | C2][2] (sum, sumsq) per row segment of the two concatenated sources.
// All reductions run in a fixed order: deterministic.
__global__ void __launch_bounds__(GN_THREADS) gn_apply_fused_k(const bf16* __restrict__ x1, int C1, int ld1,
                                                              const bf16* __restrict__ x2, int C2, int ld2, int HW,
                                                              int groups, float eps, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ film, int film_ld, int silu,
                                                              const float* __restrict__ st1, int P1,
                                                              const float* __restrict__ st2, int P2, int slabs,
                                                              bf16* __restrict__ out) {
    __shared__ float s_mean[32], s_rstd[32];
    __shared__ float2 s_ch[2048];  // per-channel (sum, sumsq) of this image
    const int C = C1 + C2;
    const int CV = C / 8;
    const int PL = GN_THREADS / CV;
    const int n = blockIdx.x, slab = blockIdx.y;
    const int pix_per_slab = HW / slabs;
    const int cpg = C / groups;
    ptx::pdl_wait();  // PDL: x and the partials come from the predecessor (the producer GEMM)
    ptx::pdl_trigger();
    // (0) start streaming: the first 4 pixel vectors of this thread are requested BEFORE the statistics prologue, whose
    //     latency (two barriers, dependent loads of the partials) then hides behind them
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const long long base = (long long)n * HW + (long long)slab * pix_per_slab;
    const bool pre = pl < PL && pl + 3 * PL < pix_per_slab;
    bf16x8 v0[4];
    if (pre) {
#pragma unroll
        for (int u = 0; u < 4; ++u) v0[u] = ld8(gn_src(x1, C1, ld1, x2, ld2, base + pl + u * PL, cv * 8));
    }
    // (1) per-channel totals over the P1 / P2 row-segment partials of this image (coalesced, fixed order)
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        const bool first = c < C1;
        const float* base = first ? st1 + ((long long)n * P1 * C1 + c) * 2 : st2 + ((long long)n * P2 * C2 + (c - C1)) * 2;
        const int P = first ? P1 : P2;
        const long long stride = (long long)(first ? C1 : C2) * 2;
        float s = 0.f, q = 0.f;
        for (int seg = 0; seg < P; ++seg) {
            const float2 v = *reinterpret_cast<const float2*>(base + seg * stride);
            s += v.x;
            q += v.y;
        }
        s_ch[c] = make_float2(s, q);
    }
    __syncthreads();
    // (2) channels -> groups: thread (g, sub) sums every 8th channel of group g, then a fixed shuffle tree
    {
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        float s = 0.f, q = 0.f;
        if (g < groups) {
            for (int i = sub; i < cpg; i += 8) {
                const float2 v = s_ch[g * cpg + i];
                s += v.x;
                q += v.y;
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && sub == 0) {
            const float cnt = (float)cpg * (float)HW;
            const float mean = s / cnt;
            const float var = fmaxf(q / cnt - mean * mean, 0.f);
            s_mean[g] = mean;
            s_rstd[g] = rsqrtf(var + eps);
        }
    }
    __syncthreads();
    if (pl >= PL) return;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j;
        const int g = c / cpg;
        float aa = s_rstd[g] * gamma[c];
        float bb = beta[c] - s_mean[g] * aa;
        if (film) {
            const float sc = 1.f + film[(long long)n * film_ld + c];
            const float sh = film[(long long)n * film_ld + C + c];
            aa *= sc;
            bb = bb * sc + sh;
        }
        a[j] = aa;
        b[j] = bb;
    }
    int p = pl;
    if (pre) {
        // the 4 vectors fetched before the statistics prologue
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8(v0[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = fmaf(f[j], a[j], b[j]);
                f[j] = silu ? silu_f(t) : t;
            }
            st8(out + (base + p + u * PL) * C + cv * 8, pack8(f));
        }
        p += 4 * PL;
    }
    for (; p + 3 * PL < pix_per_slab; p += 4 * PL) {
        bf16x8 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld8(gn_src(x1, C1, ld1, x2, ld2, base + p + u * PL, cv * 8));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = fmaf(f[j], a[j], b[j]);
                f[j] = silu ? silu_f(t) : t;
            }
            st8(out + (base + p + u * PL) * C + cv * 8, pack8(f));
        }
    }
    for (; p < pix_per_slab; p += PL) {
        float f[8];
        unpack8(ld8(gn_src(x1, C1, ld1, x2, ld2, base + p, cv * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(f[j], a[j], b[j]);
            f[j] = silu ? silu_f(t) : t;
        }
        st8(out + (base + p) * C + cv * 8, pack8(f));
    }
}

// ---- finalize + apply: statistics -> per-(image, channel) affine (a, b), then a pure streaming pass  y = act(a*x + b)
// st1 / st2 = [N][P1 | P2][C1 | C2][2] partial (sum, sumsq) from the producer GEMM epilogues. One CTA per image; every
// reduction runs in a fixed order (deterministic).
static constexpr int GN_FIN_THREADS = 1024;
__global__ void __launch_bounds__(GN_FIN_THREADS) gn_finalize_k(const float* __restrict__ st1, int P1, int C1,
                                                               const float* __restrict__ st2, int P2, int C2, int HW, int groups,
                                                               float eps, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ film,
                                                               int film_ld, float2* __restrict__ ab, float2* __restrict__ mr) {
    __shared__ float s_mean[32], s_rstd[32];
    __shared__ float2 s_ch[2048];
    __shared__ float2 s_part[2048];
    const int C = C1 + C2;
    const int n = blockIdx.x;
    const int cpg = C / groups;
    // PDL: wait for the producer of the partials, THEN let the apply kernel's CTAs be scheduled: an apply CTA that starts
    // early therefore knows the producer GEMM has completed and may prefetch x before its own wait (gn_apply_ab_k)
    ptx::pdl_wait();
    ptx::pdl_trigger();
    // per-channel totals over the P1 / P2 partials of this image (HW/16 per image with 16-row segments): `parts` thread
    // groups split a channel's partial list, then a fixed-order combine - the order depends on (C, P) only, never on the batch
    const int parts = C >= GN_FIN_THREADS ? 1 : GN_FIN_THREADS / C;
    for (int idx = threadIdx.x; idx < C * parts; idx += GN_FIN_THREADS) {
        const int c = idx % C, part = idx / C;
        const bool first = c < C1;
        const float* base = first ? st1 + ((long long)n * P1 * C1 + c) * 2 : st2 + ((long long)n * P2 * C2 + (c - C1)) * 2;
        const int P = first ? P1 : P2;
        const long long stride = (long long)(first ? C1 : C2) * 2;
        float s = 0.f, q = 0.f;
        int seg = (int)((long long)part * P / parts);
        const int seg_end = (int)((long long)(part + 1) * P / parts);
        for (; seg + 4 <= seg_end; seg += 4) {  // 4 independent loads in flight, summed in a fixed order
            const float2 v0 = *reinterpret_cast<const float2*>(base + (seg + 0) * stride);
            const float2 v1 = *reinterpret_cast<const float2*>(base + (seg + 1) * stride);
            const float2 v2 = *reinterpret_cast<const float2*>(base + (seg + 2) * stride);
            const float2 v3 = *reinterpret_cast<const float2*>(base + (seg + 3) * stride);
            s += (v0.x + v1.x) + (v2.x + v3.x);
            q += (v0.y + v1.y) + (v2.y + v3.y);
        }
        for (; seg < seg_end; ++seg) {
            const float2 v = *reinterpret_cast<const float2*>(base + seg * stride);
            s += v.x;
            q += v.y;
        }
        s_part[idx] = make_float2(s, q);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GN_FIN_THREADS) {
        float2 a = s_part[c];
        for (int part = 1; part < parts; ++part) {
            const float2 v = s_part[part * C + c];
            a.x += v.x;
            a.y += v.y;
        }
        s_ch[c] = a;
    }
    __syncthreads();
    {
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        float s = 0.f, q = 0.f;
        if (g < groups) {
            for (int i = sub; i < cpg; i += 8) {
                const float2 v = s_ch[g * cpg + i];
                s += v.x;
                q += v.y;
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && sub == 0) {
            const float cnt = (float)cpg * (float)HW;
            const float mean = s / cnt;
            const float var = fmaxf(q / cnt - mean * mean, 0.f);
            s_mean[g] = mean;
            s_rstd[g] = rsqrtf(var + eps);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GN_FIN_THREADS) {
        const int g = c / cpg;
        float aa = s_rstd[g] * gamma[c];
        float bb = beta[c] - s_mean[g] * aa;
        if (film) {
            const float sc = 1.f + film[(long long)n * film_ld + c];
            const float sh = film[(long long)n * film_ld + C + c];
            aa *= sc;
            bb = bb * sc + sh;
        }
        ab[(long long)n * C + c] = make_float2(aa, bb);
    }
    if (mr && threadIdx.x < groups) mr[(long long)n * groups + threadIdx.x] = make_float2(s_mean[threadIdx.x], s_rstd[threadIdx.x]);
}

template <int U>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_ab_k(const bf16* __restrict__ x1, int C1, int ld1,
                                                           const bf16* __restrict__ x2, int C2, int ld2, int HW,
                                                           const float2* __restrict__ ab, int silu, int slabs,
                                                           bf16* __restrict__ out) {
    const int C = C1 + C2;
    const int CV = C / 8;
    const int PL = GN_THREADS / CV;
    const int n = blockIdx.x, slab = blockIdx.y;
    const int pix_per_slab = HW / slabs;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const bool active = pl < PL;
    const long long base = (long long)n * HW + (long long)slab * pix_per_slab;
    // PDL: this kernel is only ever launched right after gn_finalize_k, which triggers its dependents AFTER its own wait - so
    // the producer of x has completed by the time any CTA of this kernel runs: the first batch of x is requested before the
    // wait for the affine table (hides gn_finalize_k's latency and the launch gap)
    int p = pl;
    bf16x8 v[U];
    const bool pre = active && p + (U - 1) * PL < pix_per_slab;
    if (pre) {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld8(gn_src(x1, C1, ld1, x2, ld2, base + p + u * PL, cv * 8));
    }
    ptx::pdl_wait();
    ptx::pdl_trigger();
    if (!active) return;
    float a[8], b[8];
    {
        const float4* src = reinterpret_cast<const float4*>(ab + (long long)n * C + cv * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = src[j];
            a[2 * j] = t.x;
            b[2 * j] = t.y;
            a[2 * j + 1] = t.z;
            b[2 * j + 1] = t.w;
        }
    }
    bool have = pre;
    for (; p + (U - 1) * PL < pix_per_slab; p += U * PL) {
        if (!have) {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld8(gn_src(x1, C1, ld1, x2, ld2, base + p + u * PL, cv * 8));
        }
        have = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = fmaf(f[j], a[j], b[j]);
                f[j] = silu ? silu_f(t) : t;
            }
            st8(out + (base + p + u * PL) * C + cv * 8, pack8(f));
        }
    }
    for (; p < pix_per_slab; p += PL) {
        float f[8];
        unpack8(ld8(gn_src(x1, C1, ld1, x2, ld2, base + p, cv * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(f[j], a[j], b[j]);
            f[j] = silu ? silu_f(t) : t;
        }
        st8(out + (base + p) * C + cv * 8, pack8(f));
    }
}

static int g_gn_unroll = 4;
void set_gn_unroll(int v) { g_gn_unroll = v; }

int gn_apply_slabs(int N, int HW, int C) {
    // enough CTAs for ~16 per SM, but keep >= 4 pixels per pixel-lane per slab so the unrolled loop has work
    const int PL = GN_THREADS / (C / 8) > 0 ? GN_THREADS / (C / 8) : 1;
    int slabs = 1;
    while (N * slabs < 2368 && HW / (slabs * 2) >= 8 * PL && HW % (slabs * 2) == 0) slabs *= 2;
    return slabs;
}

void gn_apply_fused(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups, float eps,
                    const float* gamma, const float* beta, const float* film, int film_ld, int silu, const float* st1,
                    int P1, const float* st2, int P2, bf16* out, cudaStream_t st) {
    const int slabs = gn_apply_slabs(N, HW, C1 + C2);
    dim3 grid(N, slabs);
    launch_pdl(gn_apply_fused_k, grid, dim3(GN_THREADS), st, x1, C1, ld1, x2, C2, ld2, HW, groups, eps, gamma, beta, film, film_ld, silu, st1, P1,
               st2, P2, slabs, out);
}

void gn_finalize_apply(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups,
                       float eps, const float* gamma, const float* beta, const float* film, int film_ld, int silu,
                       const float* st1, int P1, const float* st2, int P2, float* ab_ws, bf16* out, cudaStream_t st, float* mr) {
    launch_pdl(gn_finalize_k, dim3(N), dim3(GN_FIN_THREADS), st, st1, P1, C1, st2, P2, C2, HW, groups, eps, gamma, beta, film, film_ld,
               reinterpret_cast<float2*>(ab_ws), reinterpret_cast<float2*>(mr));
    const int slabs = gn_apply_slabs(N, HW, C1 + C2);
    dim3 grid(N, slabs);
    if (g_gn_unroll == 8)
        launch_pdl(gn_apply_ab_k<8>, grid, dim3(GN_THREADS), st, x1, C1, ld1, x2, C2, ld2, HW, reinterpret_cast<const float2*>(ab_ws), silu, slabs, out);
    else
        launch_pdl(gn_apply_ab_k<4>, grid, dim3(GN_THREADS), st, x1, C1, ld1, x2, C2, ld2, HW, reinterpret_cast<const float2*>(ab_ws), silu, slabs, out);
}

// [N][P][C][2] -> [N][1][C][2]: collapses many row-segment partials (large feature maps) so that the apply kernel's
// per-CTA prologue stays small.  Fixed summation order.
__global__ void gn_collapse_k(const float* __restrict__ in, float* __restrict__ out, int P, int C) {
    const int n = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float* base = in + ((long long)n * P * C + c) * 2;
    float s = 0.f, q = 0.f;
    for (int seg = 0; seg < P; ++seg) {
        const float2 v = *reinterpret_cast<const float2*>(base + (long long)seg * C * 2);
        s += v.x;
        q += v.y;
    }
    *reinterpret_cast<float2*>(out + ((long long)n * C + c) * 2) = make_float2(s, q);
}
void gn_collapse(const float* in, float* out, int N, int P, int C, cudaStream_t st) {
    dim3 grid((C + 127) / 128, N);
    gn_collapse_k<<<grid, 128, 0, st>>>(in, out, P, C);
}

__global__ void silu_to_bf16_k(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(silu_f(x[i]));
}
void silu_to_bf16(const float* x, bf16* y, long long n, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    silu_to_bf16_k<<<(int)blocks, 256, 0, st>>>(x, y, n);
}

// ============================================================================================ small dense helpers

__global__ void timestep_embedding_k(const float* __restrict__ t, float* __restrict__ out, int N, int dim, int order) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * half) return;
    const int n = i / half, j = i % half;
    // same fp32 arithmetic as the reference: exp(arange * -(ln 1e4 / denom)) then t * freq
    float freq;
    if (order == 0) {
        const float e = logf(10000.f) / (float)(half - 1);
        freq = expf((float)j * -e);
    } else {
        freq = expf(-logf(10000.f) * (float)j / (float)half);
    }
    const float a = t[n] * freq;
    const float s = sinf(a), c = cosf(a);
    if (order == 0) {
        out[(long long)n * dim + j] = s;
        out[(long long)n * dim + half + j] = c;
    } else {
        out[(long long)n * dim + j] = c;
        out[(long long)n * dim + half + j] = s;
    }
}
void timestep_embedding(const float* t, float* out, int N, int dim, int order, cudaStream_t st) {
    const int total = N * (dim / 2);
    timestep_embedding_k<<<(total + 127) / 128, 128, 0, st>>>(t, out, N, dim, order);
}

static constexpr int LIN_ROWS = 16;
// block = 8 warps; each warp produces one output column for LIN_ROWS rows; x tile staged (activated) in smem.
__global__ void linear_f32_k(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                             const float* __restrict__ b, float* __restrict__ y, int ldy, int N, int K, int O,
                             int act_in, int act_out) {
    extern __shared__ float sx[];  // [LIN_ROWS][K]
    const int row0 = blockIdx.y * LIN_ROWS;
    for (int i = threadIdx.x; i < LIN_ROWS * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        float v = (row0 + r < N) ? x[(long long)(row0 + r) * ldx + k] : 0.f;
        sx[i] = act_f(v, act_in);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= O) return;
    float acc[LIN_ROWS];
#pragma unroll
    for (int r = 0; r < LIN_ROWS; ++r) acc[r] = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float wv = W[(long long)o * K + k];
#pragma unroll
        for (int r = 0; r < LIN_ROWS; ++r) acc[r] = fmaf(wv, sx[r * K + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < LIN_ROWS; ++r) {
        const float s = warp_sum(acc[r]);
        if (lane == 0 && row0 + r < N) y[(long long)(row0 + r) * ldy + o] = act_f(s + (b ? b[o] : 0.f), act_out);
    }
}
void linear_f32(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int N, int K, int O,
                int act_in, int act_out, cudaStream_t st) {
    dim3 grid((O + 7) / 8, (N + LIN_ROWS - 1) / LIN_ROWS);
    const size_t smem = (size_t)LIN_ROWS * K * sizeof(float);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(linear_f32_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        configured.set();
    }
    linear_f32_k<<<grid, 256, smem, st>>>(x, ldx, W, b, y, ldy, N, K, O, act_in, act_out);
}

__global__ void embedding_add_k(float* __restrict__ emb, const float* __restrict__ table,
                                const long long* __restrict__ idx, int N, int D) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)N * D) return;
    const int n = (int)(i / D), d = (int)(i % D);
    emb[i] += table[idx[n] * D + d];
}
void embedding_add(float* emb, const float* table, const long long* idx, int N, int D, cudaStream_t st) {
    const long long total = (long long)N * D;
    embedding_add_k<<<(int)((total + 255) / 256), 256, 0, st>>>(emb, table, idx, N, D);
}

// ============================================================================================ resampling

__global__ void upsample2x_k(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int CV) {
    const long long total = (long long)N * (2 * H) * (2 * W) * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int ow = (int)(r % (2 * W));
        r /= (2 * W);
        const int oh = (int)(r % (2 * H));
        const int n = (int)(r / (2 * H));
        const bf16x8 v = ld8(x + ((((long long)n * H + (oh >> 1)) * W + (ow >> 1)) * CV + cv) * 8);
        st8(out + i * 8, v);
    }
}
__global__ void nhwc_to_nchw_f32_k(const float* __restrict__ src, int ld, float* __restrict__ out, long long total, int HW, int C) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / HW;
        const int p = (int)(i - n * HW);
        const float4 v = *reinterpret_cast<const float4*>(src + i * ld);  // ld % 4 == 0, C <= 4
        float* o = out + n * C * HW + p;
        o[0] = v.x;
        if (C > 1) o[HW] = v.y;
        if (C > 2) o[2LL * HW] = v.z;
        if (C > 3) o[3LL * HW] = v.w;
    }
}
void nhwc_to_nchw_f32(const float* src, int ld, float* out, int N, int HW, int C, cudaStream_t st) {
    const long long total = (long long)N * HW;
    long long blocks = (total + 255) / 256;
    if (blocks > 2368) blocks = 2368;
    nhwc_to_nchw_f32_k<<<(int)blocks, 256, 0, st>>>(src, ld, out, total, HW, C);
}

void upsample2x(const bf16* x, bf16* out, int N, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)N * 4 * H * W * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    upsample2x_k<<<(int)blocks, 256, 0, st>>>(x, out, N, H, W, C / 8);
}

__global__ void avgpool2_k(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int CV, int act) {
    const int OH = H / 2, OW = W / 2;
    const long long total = (long long)N * OH * OW * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int ow = (int)(r % OW);
        r /= OW;
        const int oh = (int)(r % OH);
        const int n = (int)(r / OH);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float f[8];
                unpack8(ld8(x + ((((long long)n * H + 2 * oh + dy) * W + 2 * ow + dx) * CV + cv) * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = act_f(acc[j] * 0.25f, act);
        st8(out + i * 8, pack8(acc));
    }
}
void avgpool2(const bf16* x, bf16* out, int N, int H, int W, int C, int act, cudaStream_t st) {
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    avgpool2_k<<<(int)blocks, 256, 0, st>>>(x, out, N, H, W, C / 8, act);
}

// ============================================================================================ tiny attention

// one CTA per (image, head); q/k/v tiles in smem as bf16, scores fp32. seq <= 64, d % 8 == 0, ld % 8 == 0.
// Rows are padded to d + 2 elements: the pitch in 32-bit words (d/2 + 1) is odd, so the 16 key rows a warp walks in the score
// loop fall into distinct banks (the unpadded layout was a 16-way bank conflict: 57 us per launch for a 16-token map at B = 256).
// Global loads are 16-byte vectors; every dot product walks bf16 pairs.
__global__ void __launch_bounds__(256) attn_small_k(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int ld,
                                                    bf16* __restrict__ out, int ldo, int heads, int seq, int d, float scale) {
    extern __shared__ __align__(16) uint8_t smraw[];
    const int dp = d + 2;
    bf16* sq = reinterpret_cast<bf16*>(smraw);
    bf16* sk = sq + seq * dp;
    bf16* sv = sk + seq * dp;
    float* ss = reinterpret_cast<float*>(sv + seq * dp);  // [seq][seq]
    const int n = blockIdx.x / heads, h = blockIdx.x % heads;
    const long long base = (long long)n * seq * ld + (long long)h * d;
    const int dv = d >> 3;
    for (int i = threadIdx.x; i < seq * dv; i += blockDim.x) {
        const int t = i / dv, c = (i - t * dv) * 8;
        const long long g = base + (long long)t * ld + c;
        const uint4 a = *reinterpret_cast<const uint4*>(q + g);
        const uint4 b = *reinterpret_cast<const uint4*>(k + g);
        const uint4 e = *reinterpret_cast<const uint4*>(v + g);
        uint32_t* dq = reinterpret_cast<uint32_t*>(sq + t * dp + c);
        uint32_t* dk = reinterpret_cast<uint32_t*>(sk + t * dp + c);
        uint32_t* dw = reinterpret_cast<uint32_t*>(sv + t * dp + c);
        dq[0] = a.x; dq[1] = a.y; dq[2] = a.z; dq[3] = a.w;
        dk[0] = b.x; dk[1] = b.y; dk[2] = b.z; dk[3] = b.w;
        dw[0] = e.x; dw[1] = e.y; dw[2] = e.z; dw[3] = e.w;
    }
    __syncthreads();
    const int d2 = d >> 1;
    for (int i = threadIdx.x; i < seq * seq; i += blockDim.x) {
        const int a = i / seq, b = i - a * seq;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(sq + a * dp);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(sk + b * dp);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
        for (int c = 0; c < d2; ++c) {
            const float2 x = __bfloat1622float2(pa[c]), y = __bfloat1622float2(pb[c]);
            s0 = fmaf(x.x, y.x, s0);
            s1 = fmaf(x.y, y.y, s1);
        }
        ss[i] = (s0 + s1) * scale;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int a = threadIdx.x >> 5; a < seq; a += blockDim.x >> 5) {
        float m = -CUDART_INF_F;
        for (int b = lane; b < seq; b += 32) m = fmaxf(m, ss[a * seq + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float l = 0.f;
        for (int b = lane; b < seq; b += 32) {
            const float e = __expf(ss[a * seq + b] - m);
            ss[a * seq + b] = e;
            l += e;
        }
        l = warp_sum(l);
        const float inv = 1.f / l;
        for (int b = lane; b < seq; b += 32) ss[a * seq + b] *= inv;
    }
    __syncthreads();
    const long long obase = (long long)n * seq * ldo + (long long)h * d;
    for (int i = threadIdx.x; i < seq * d2; i += blockDim.x) {
        const int a = i / d2, c = i - a * d2;
        float s0 = 0.f, s1 = 0.f;
        for (int b = 0; b < seq; ++b) {
            const float p = ss[a * seq + b];
            const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + b * dp + 2 * c));
            s0 = fmaf(p, y.x, s0);
            s1 = fmaf(p, y.y, s1);
        }
        *reinterpret_cast<__nv_bfloat162*>(out + obase + (long long)a * ldo + 2 * c) = __floats2bfloat162_rn(s0, s1);
    }
}
void attn_small(const bf16* q, const bf16* k, const bf16* v, int ld, bf16* out, int ldo, int N, int heads, int seq,
                int d, float scale, cudaStream_t st) {
    const size_t smem = (size_t)3 * seq * (d + 2) * sizeof(bf16) + (size_t)seq * seq * sizeof(float);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(attn_small_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        configured.set();
    }
    attn_small_k<<<N * heads, 256, smem, st>>>(q, k, v, ld, out, ldo, heads, seq, d, scale);
}

// ============================================================================================ sampler transitions

// fp32 sample in [-1, 1] -> uint8: ((x + 1) * 127.5).clamp(0, 255) truncated (generate_large.py:43, generate_cifar10.py:205-209);
// the same expression as quantize_u8_k, so the fused and the stand-alone forms agree bit for bit
__device__ __forceinline__ uint8_t quant1(float x) {
    float v = (x + 1.f) * 127.5f;
    v = fminf(fmaxf(v, 0.f), 255.f);
    return (uint8_t)v;
}
__device__ __forceinline__ uchar4 quant4(const float (&v)[4]) { return make_uchar4(quant1(v[0]), quant1(v[1]), quant1(v[2]), quant1(v[3])); }

// one CTA per sample (logp is a per-sample reduction).
__global__ void var_step_k(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z,
                           const float* __restrict__ a, const float* __restrict__ c, const float* __restrict__ sigma,
                           float* __restrict__ xn, float* __restrict__ mean, float* __restrict__ control,
                           float* __restrict__ logp, uint8_t* __restrict__ u8, int CHW) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    const float an = a[n], cn = c[n], sn = sigma[n];
    const float inv2var = 1.f / (2.f * sn * sn);
    const float cst = -logf(sn) - 0.9189385332046727f;  // -ln(sigma) - 0.5 ln(2 pi)
    const long long base = (long long)n * CHW;
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < CHW; i += blockDim.x * 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + base + i);
        const float4 ev = *reinterpret_cast<const float4*>(eps + base + i);
        const float4 zv = *reinterpret_cast<const float4*>(z + base + i);
        float xs[4] = {xv.x, xv.y, xv.z, xv.w}, es[4] = {ev.x, ev.y, ev.z, ev.w}, zs[4] = {zv.x, zv.y, zv.z, zv.w};
        float ctl[4], mu[4], nx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xm = xs[j] * an;       // x *= x_prev_multiplier
            ctl[j] = cn * es[j];               // control = c * eps
            mu[j] = xm + ctl[j];               // pred_mean
            nx[j] = xm + (ctl[j] + sn * zs[j]);  // x += control + sigma * z   (reference association order)
            const float dlt = nx[j] - mu[j];
            acc += -(dlt * dlt) * inv2var + cst;
        }
        *reinterpret_cast<float4*>(xn + base + i) = make_float4(nx[0], nx[1], nx[2], nx[3]);
        if (u8) *reinterpret_cast<uchar4*>(u8 + base + i) = quant4(nx);
        if (mean) *reinterpret_cast<float4*>(mean + base + i) = make_float4(mu[0], mu[1], mu[2], mu[3]);
        if (control) *reinterpret_cast<float4*>(control + base + i) = make_float4(ctl[0], ctl[1], ctl[2], ctl[3]);
    }
    const float tot = block_sum(acc, sh);
    if (threadIdx.x == 0 && logp) logp[n] = tot / (float)CHW;
}
void var_step(const float* x, const float* eps, const float* z, const float* a, const float* c, const float* sigma,
              float* xn, float* mean, float* control, float* logp, uint8_t* u8, int N, int CHW, cudaStream_t st) {
    var_step_k<<<N, 256, 0, st>>>(x, eps, z, a, c, sigma, xn, mean, control, logp, u8, CHW);
}

__global__ void edm_step_k(const float* __restrict__ x, const float* __restrict__ F, const float* __restrict__ z,
                           const float* __restrict__ coef, float* __restrict__ xn, float* __restrict__ mean,
                           uint8_t* __restrict__ u8, int CHW4, long long total4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / CHW4);
        const float c_skip = coef[n * 5 + 0], c_out = coef[n * 5 + 1], sg = coef[n * 5 + 2], sd = coef[n * 5 + 3],
                    sn = coef[n * 5 + 4];
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        const float4 fv = reinterpret_cast<const float4*>(F)[i];
        const float4 zv = reinterpret_cast<const float4*>(z)[i];
        float xs[4] = {xv.x, xv.y, xv.z, xv.w}, fs[4] = {fv.x, fv.y, fv.z, fv.w}, zs[4] = {zv.x, zv.y, zv.z, zv.w};
        float mu[4], nx[4];
        const float dt = sd - sg;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float D = c_out * fs[j] + c_skip * xs[j];
            const float dd = (xs[j] - D) / sg;
            mu[j] = xs[j] + dd * dt;
            nx[j] = mu[j] + zs[j] * sn;
        }
        reinterpret_cast<float4*>(xn)[i] = make_float4(nx[0], nx[1], nx[2], nx[3]);
        if (u8) reinterpret_cast<uchar4*>(u8)[i] = quant4(nx);
        if (mean) reinterpret_cast<float4*>(mean)[i] = make_float4(mu[0], mu[1], mu[2], mu[3]);
    }
}
void edm_step(const float* x, const float* F, const float* z, const float* coef, float* xn, float* mean, uint8_t* u8, int N,
              int CHW, cudaStream_t st) {
    const long long total4 = (long long)N * CHW / 4;
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    edm_step_k<<<(int)blocks, 256, 0, st>>>(x, F, z, coef, xn, mean, u8, CHW / 4, total4);
}

// per-step coefficient broadcast for the EDM rollout: coef[n] = v[2..6], x_scale[n] = v[0], t[n] = v[1]
struct Coef6 {
    float v[6];
};
__global__ void edm_fill_k(float* __restrict__ coef, float* __restrict__ x_scale, float* __restrict__ t, int N, Coef6 c,
                           const float* __restrict__ sigma_noise) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    x_scale[n] = c.v[0];
    t[n] = c.v[1];
#pragma unroll
    for (int j = 0; j < 4; ++j) coef[n * 5 + j] = c.v[2 + j];
    coef[n * 5 + 4] = *sigma_noise;
}
void edm_fill(float* coef, float* x_scale, float* t, int N, const float* row6, const float* sigma_noise_dev, cudaStream_t st) {
    Coef6 c;
    for (int j = 0; j < 6; ++j) c.v[j] = row6[j];
    edm_fill_k<<<(N + 127) / 128, 128, 0, st>>>(coef, x_scale, t, N, c, sigma_noise_dev);
}

// per-step coefficient broadcast for the VAR rollout: t[n] = tau, a[n] = a, c[n] = c (host scalars), sigma[n] = *sigma_dev
__global__ void var_fill_k(float* __restrict__ t, float* __restrict__ a, float* __restrict__ c, float* __restrict__ sigma, int N,
                           float tau, float av, float cv, const float* __restrict__ sigma_dev) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    t[n] = tau;
    a[n] = av;
    c[n] = cv;
    sigma[n] = *sigma_dev;
}
void var_fill(float* t, float* a, float* c, float* sigma, int N, float tau, float av, float cv, const float* sigma_dev,
              cudaStream_t st) {
    var_fill_k<<<(N + 127) / 128, 128, 0, st>>>(t, a, c, sigma, N, tau, av, cv, sigma_dev);
}

// ============================================================================================ value head

__global__ void value_head_k(const bf16* __restrict__ h, int HW, int C, const float* __restrict__ lin_w,
                             const float* __restrict__ lin_b, const float* __restrict__ scale_w,
                             const float* __restrict__ scale_b, float* __restrict__ out) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += fmaxf(__bfloat162float(h[((long long)n * HW + p) * C + c]), 0.f);
        acc = fmaf(s, lin_w[c], acc);
    }
    const float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        float v = tot + lin_b[0];
        if (scale_w) v = v * scale_w[0] + scale_b[0];
        out[n] = v;
    }
}
void value_head(const bf16* h, int N, int HW, int C, const float* lin_w, const float* lin_b, const float* scale_w,
                const float* scale_b, float* out, cudaStream_t st) {
    value_head_k<<<N, 256, 0, st>>>(h, HW, C, lin_w, lin_b, scale_w, scale_b, out);
}

// ============================================================================================ tiny utilities

__global__ void fill_f32_k(float* p, float v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
void fill_f32(float* p, float v, long long n, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    fill_f32_k<<<(int)blocks, 256, 0, st>>>(p, v, n);
}
__global__ void vec_add_f32_k(const float* a, const float* b, float* out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] + b[i];
}
void vec_add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    vec_add_f32_k<<<(int)blocks, 256, 0, st>>>(a, b, out, n);
}

// ============================================================================================ layout helpers

__global__ void nhwc_bf16_to_nchw_f32_k(const bf16* __restrict__ x, float* __restrict__ out, int C, int HW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        long long r = i / HW;
        const int c = (int)(r % C);
        const long long n = r / C;
        out[i] = __bfloat162float(x[(n * HW + p) * C + c]);
    }
}
void nhwc_bf16_to_nchw_f32(const bf16* x, float* out, int N, int C, int HW, cudaStream_t st) {
    const long long total = (long long)N * C * HW;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    nhwc_bf16_to_nchw_f32_k<<<(int)blocks, 256, 0, st>>>(x, out, C, HW, total);
}
__global__ void nchw_f32_to_nhwc_bf16_k(const float* __restrict__ x, bf16* __restrict__ out, int C, int HW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int p = (int)(r % HW);
        const long long n = r / HW;
        out[i] = __float2bfloat16_rn(x[(n * C + c) * HW + p]);
    }
}
void nchw_f32_to_nhwc_bf16(const float* x, bf16* out, int N, int C, int HW, cudaStream_t st) {
    const long long total = (long long)N * C * HW;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    nchw_f32_to_nhwc_bf16_k<<<(int)blocks, 256, 0, st>>>(x, out, C, HW, total);
}

__global__ void quantize_u8_k(const float* __restrict__ x, uint8_t* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = quant1(x[i]);  // truncation == torch .to(uint8)
}
void quantize_u8(const float* x, uint8_t* out, long long n, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    quantize_u8_k<<<(int)blocks, 256, 0, st>>>(x, out, n);
}

}  // namespace dxmi
