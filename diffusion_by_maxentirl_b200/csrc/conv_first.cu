// First convolution of a network on the inference path: 3x3, stride 1, pad 1, 3 -> Cout channels, fp32 NCHW in (optionally scaled per
// image: the EDM c_in), bf16 NHWC out + the GroupNorm partial statistics of the output (unet_small.py:262 conv_in, cm/unet.py:577-583
// input_blocks[0], modules.py:142-147 conv1 with its leaky-relu).
//
// Why a second kernel next to conv3x3_first_k (kernels.cu): that one contracts K = 27 on the FMA pipe - 3456 FFMA per warp and 32
// pixels, 75 us per launch at 32x32, B = 256 (ncu: FMA pipe 43 %, issue slots 61 %: bound by instruction issue, 5x off the 67 MB
// store stream it should be).  Here the contraction is 2 x (Cout / 8) warp-level mma.sync.m16n8k16 per 16 pixels: K = 27 zero-padded to
// 32, the im2col A fragments gathered straight from the fp32 input patch in shared memory (rounded to bf16 like every other activation of
// the path), the bf16 weights as [Cout][32 + 8] rows.  mma.sync, not tcgen05: a K = 32 contraction has nothing to pipeline and the
// accumulator (32 pixels x 32 channels per pass) lives in 32 registers.  The training plans keep the fp32 FMA kernel.
//
// One CTA = 4 warps = one 128-pixel tile (th rows x tw columns) per step, persistent over tiles; the next tile's input patch is fetched
// into registers while the current one is computed.  Output: fragments -> per-warp staging rows in shared memory -> full 16-byte NHWC
// vectors; statistics: one (sum, sum of squares) per tile and channel over the bf16-rounded outputs, summed in pixel order (fixed
// order: deterministic, bitwise batch invariant) - the layout conv3x3_first_k and the GEMM epilogues publish.
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace dxmi {

namespace {

__device__ __forceinline__ float act_first(float v, int act) {
    if (act == 1) return v > 0.f ? v : 0.2f * v;
    if (act == 2) return v / (1.f + __expf(-v));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int KP = 40;    // weight row pitch (bf16): 32 + 8 -> the 8 rows of a B fragment fall into distinct banks
constexpr int NPRE = 10;  // input patch elements per thread held in flight: >= ceil(3 * (th+2) * (tw+2) / 128) = 5 / 7 / 10 for tw = 32 / 64 / 128

}  // namespace

// grid: persistent (<= ntiles CTAs), block 128.  th * tw == 128, tw | W, th | H, Cout % 32 == 0.
__global__ void __launch_bounds__(128) conv3x3_first_tc_k(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ w,
                                                         const float* __restrict__ b, bf16* __restrict__ out, float* __restrict__ stats, int H, int W,
                                                         int Cout, int act, int th, int tw, int ntiles) {
    extern __shared__ __align__(16) uint8_t smraw[];
    const int pw = tw + 2, ph = th + 2;
    const int npatch = 3 * ph * pw;
    const int patch = (npatch + 3) & ~3;
    const int spitch = Cout * 2 + 16;  // staging row pitch (bytes): rows 4 banks apart -> conflict-free fragment writes
    bf16* sw = reinterpret_cast<bf16*>(smraw);                                  // [Cout][KP]
    float* sb = reinterpret_cast<float*>(smraw + (size_t)Cout * KP * 2);        // [Cout]
    float* sx = sb + Cout;                                                      // [2][patch]: ci-major, (th+2) x (tw+2) zero-padded
    uint8_t* stg = reinterpret_cast<uint8_t*>(sx + 2 * patch);                  // [128 pixels][spitch]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tiles_w = W / tw, tiles_per_img = tiles_w * (H / th);
    const long long HW = (long long)H * W;

    // weights: OIHW fp32 -> [co][k = tap * 3 + ci] bf16, k >= 27 zero
    for (int i = tid; i < Cout * 32; i += 128) {
        const int co = i >> 5, k = i & 31;
        float v = 0.f;
        if (k < 27) {
            const int tap = k / 3, ci = k - tap * 3;
            v = w[((long long)co * 3 + ci) * 9 + tap];
        }
        sw[co * KP + k] = __float2bfloat16_rn(v);
    }
    for (int i = tid; i < Cout; i += 128) sb[i] = b ? b[i] : 0.f;

    // im2col geometry of this thread's fragments: k indices {2t, 2t+1, 2t+8, 2t+9} (+16) -> offsets into the patch
    int koff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = (j >> 2) * 16 + ((j >> 1) & 1) * 8 + 2 * t + (j & 1);
        const int tap = k / 3, ci = k - tap * 3;
        koff[j] = k < 27 ? (ci * ph + tap / 3) * pw + tap % 3 : -1;
    }
    int pbase[4];  // [m tile][row g | g + 8]: top-left patch element of the pixel
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = warp * 32 + (j >> 1) * 16 + (j & 1) * 8 + g;
        pbase[j] = (p / tw) * pw + p % tw;
    }

    float nx[NPRE];
    auto fetch = [&](int tile) {
        const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
        const int h0 = (trem / tiles_w) * th, w0 = (trem % tiles_w) * tw;
        const float sc = in_scale ? in_scale[n] : 1.f;
#pragma unroll
        for (int u = 0; u < NPRE; ++u) {
            const int i = tid + u * 128;
            float v = 0.f;
            if (i < npatch) {
                const int ci = i / (ph * pw), r = (i / pw) % ph, c = i % pw;
                const int hh = h0 + r - 1, ww = w0 + c - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((long long)n * 3 + ci) * HW + (long long)hh * W + ww] * sc;
            }
            nx[u] = v;
        }
    };
    auto park = [&](float* dst) {
#pragma unroll
        for (int u = 0; u < NPRE; ++u) {
            const int i = tid + u * 128;
            if (i < npatch) dst[i] = nx[u];
        }
    };
    if ((int)blockIdx.x < ntiles) {
        fetch(blockIdx.x);
        park(sx);
    }
    int buf = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
        const int h0 = (trem / tiles_w) * th, w0 = (trem % tiles_w) * tw;
        __syncthreads();  // this tile's patch is parked (and the weights staged); the previous tile's staging rows have been read
        const float* sxc = sx + buf * patch;
        const bool more = tile + (int)gridDim.x < ntiles;
        if (more) fetch(tile + gridDim.x);
        // ---- A fragments: 2 m tiles x 2 k steps
        uint32_t a[2][2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // register q: row (q & 1 ? g + 8 : g), k pair (q >> 1 ? +8 : +0)
                    const float* px = sxc + pbase[m * 2 + (q & 1)];
                    const int j0 = s * 4 + (q >> 1) * 2;
                    const float v0 = koff[j0] >= 0 ? px[koff[j0]] : 0.f;
                    const float v1 = koff[j0 + 1] >= 0 ? px[koff[j0 + 1]] : 0.f;
                    a[m][s][q] = pack_bf16(v0, v1);
                }
        // ---- 32 output channels per pass
        uint8_t* wst = stg + (size_t)warp * 32 * spitch;
        for (int c0 = 0; c0 < Cout; c0 += 32) {
            float acc[2][4][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t* wr = reinterpret_cast<const uint32_t*>(sw + (c0 + j * 8 + g) * KP + 2 * t);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const uint32_t b0 = wr[s * 8], b1 = wr[s * 8 + 4];
                    mma_16816(acc[0][j], a[0][s], b0, b1);
                    mma_16816(acc[1][j], a[1][s], b0, b1);
                }
            }
            // fragment (rows g / g + 8, channels c0 + 8 j + 2 t, + 1) -> bias, activation, bf16 -> staging rows
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j * 8 + 2 * t;
                const float2 bb = *reinterpret_cast<const float2*>(sb + c);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint32_t lo = pack_bf16(act_first(acc[m][j][0] + bb.x, act), act_first(acc[m][j][1] + bb.y, act));
                    const uint32_t hi = pack_bf16(act_first(acc[m][j][2] + bb.x, act), act_first(acc[m][j][3] + bb.y, act));
                    *reinterpret_cast<uint32_t*>(wst + (size_t)(m * 16 + g) * spitch + c * 2) = lo;
                    *reinterpret_cast<uint32_t*>(wst + (size_t)(m * 16 + g + 8) * spitch + c * 2) = hi;
                }
            }
        }
        __syncwarp();
        // ---- this warp's 32 pixels -> NHWC lines, 16 bytes per lane
        {
            const int cv = Cout >> 3;  // 16-byte vectors per pixel
            const int total = 32 * cv;
            for (int v = lane; v < total; v += 32) {
                const int pl = v / cv, c8 = v - pl * cv;
                const int p = warp * 32 + pl;
                const int pr = p / tw, pc = p - pr * tw;
                const uint4 u = *reinterpret_cast<const uint4*>(wst + (size_t)pl * spitch + c8 * 16);
                *reinterpret_cast<uint4*>(out + ((long long)n * HW + (long long)(h0 + pr) * W + w0 + pc) * Cout + c8 * 8) = u;
            }
        }
        if (stats) {
            __syncthreads();  // all four warps' rows are staged
            for (int c2 = tid; c2 < (Cout >> 1); c2 += 128) {
                // four interleaved chains (pixel index mod 4), combined in a fixed order: a single chain is 128 dependent adds
                float s0[4] = {0.f, 0.f, 0.f, 0.f}, q0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, q1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                for (int p = 0; p < 128; p += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(stg + (size_t)(p + u) * spitch + c2 * 4));
                        s0[u] += v.x;
                        q0[u] = fmaf(v.x, v.x, q0[u]);
                        s1[u] += v.y;
                        q1[u] = fmaf(v.y, v.y, q1[u]);
                    }
                }
                *reinterpret_cast<float4*>(stats + ((long long)tile * Cout + 2 * c2) * 2) =
                    make_float4((s0[0] + s0[1]) + (s0[2] + s0[3]), (q0[0] + q0[1]) + (q0[2] + q0[3]), (s1[0] + s1[1]) + (s1[2] + s1[3]),
                                (q1[0] + q1[1]) + (q1[2] + q1[3]));
            }
        }
        // the next tile's patch (requested at the top of this iteration) goes into the other buffer: its last readers finished before this
        // iteration's first barrier, and the next iteration's barrier orders these writes before its reads
        if (more) park(sx + (buf ^ 1) * patch);
    }
}

static size_t conv_first_tc_smem(int W, int Cout) {
    const int tw = W < 128 ? W : 128, th = 128 / tw;
    const int patch = (3 * (th + 2) * (tw + 2) + 3) & ~3;
    return (size_t)Cout * KP * 2 + (size_t)Cout * 4 + (size_t)2 * patch * 4 + (size_t)128 * (Cout * 2 + 16);
}

bool conv3x3_first_tc_supported(int Cin, int H, int W, int Cout) {
    if (Cin != 3 || Cout % 32 || Cout > 512 || (H * W) % 128 || (W & (W - 1)) || W < 8) return false;
    const int tw = W < 128 ? W : 128, th = 128 / tw;
    if (H % th || 3 * (th + 2) * (tw + 2) > NPRE * 128) return false;
    return conv_first_tc_smem(W, Cout) <= 160 * 1024;
}

void conv3x3_first_tc(const float* x, const float* in_scale, const float* w, const float* b, bf16* out, float* stats, int N, int H, int W,
                      int Cout, int act, cudaStream_t st) {
    const int tw = W < 128 ? W : 128, th = 128 / tw;
    const size_t smem = conv_first_tc_smem(W, Cout);
    static DevFlags configured;
    static int num_sms = 148;
    if (!configured.test()) {
        cudaFuncSetAttribute(conv3x3_first_tc_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.set();
    }
    const int ntiles = N * (H / th) * (W / tw);
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int blocks = num_sms * per_sm;
    if (blocks > ntiles) blocks = ntiles;
    conv3x3_first_tc_k<<<blocks, 128, smem, st>>>(x, in_scale, w, b, out, stats, H, W, Cout, act, th, tw, ntiles);
}

}  // namespace dxmi
