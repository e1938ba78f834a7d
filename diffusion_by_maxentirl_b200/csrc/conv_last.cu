// Last convolution of a U-Net: 3x3, stride 1, pad 1, C -> Cout <= 8 channels (3 in every DxMI config), NHWC bf16 in, fp32 NCHW out
// (unet_small.py:330-332 conv_out, cm/unet.py:738-742 `out`).
//
// Why not the tcgen05 GEMM: with 3 output channels the contraction is 0.03 % of a U-Net's FLOPs but the implicit-GEMM kernels
// pay for it like a 32-column tile AND re-read the input once per tap through L2 (9 x 67 MB at CIFAR B = 256: 74-92 us).  The
// operation is a read-once HBM stream: this kernel stages a (rows + 2) x (W + 2) x C halo tile of the input in shared memory ONCE,
// keeps the 8 x 9C weights next to it, and contracts with warp-level mma.sync.m16n8k16 - the one place where the legacy tensor-core
// instruction is the right tool: its N = 8 tile is exactly the (zero-padded) output width, there is no accumulator to drain and
// no epilogue worth overlapping.  One CTA = 128 output pixels (TH full image rows), one warp = 32 pixels = two m16 tiles (each inside one image row: 16 | W).
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace dxmi {

namespace {

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}

}  // namespace

// x [N,H,W,C] bf16; wp [8][9*C] bf16 (row co, k = tap*C + c, rows >= Cout zero); bias8 [8] fp32; out [N,Cout,H,W] fp32.
// grid = N * H / TH, block = 128, TH * W == 128.
__global__ void __launch_bounds__(128) conv3x3_last_k(const bf16* __restrict__ x, const bf16* __restrict__ wp, const float* __restrict__ bias8,
                                                      float* __restrict__ out, int H, int W, int C, int Cout, int TH) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int CP = C + 8;            // padded pixel pitch (elements): 16 extra bytes -> conflict-free ldmatrix rows
    const int KP = 9 * C + 8;        // padded weight-row pitch
    bf16* st = reinterpret_cast<bf16*>(sm);                       // [(TH+2)][W+2][CP]
    bf16* sw = st + (size_t)(TH + 2) * (W + 2) * CP;              // [8][KP]
    const int tiles_per_img = H / TH;
    const int n = blockIdx.x / tiles_per_img;
    const int y0 = (blockIdx.x - n * tiles_per_img) * TH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- stage the halo tile (zero outside the image) and the weights: 16-byte vectors
    const int CV = C / 8;
    const int tile_vecs = (TH + 2) * (W + 2) * CV;
    // (cp.async: every thread issues ALL of its 16-byte copies back to back - ~50 KB in flight per CTA.  The first version staged through
    // registers, one dependent load per loop trip: 6 KB in flight per SM, 102 us for a 67 MB read-once stream)
    // (index arithmetic is incremental: the three runtime integer divisions per 16-byte copy of the first version cost more
    //  issue slots than everything else in the kernel - 56 us of a launch that should be a 67 MB stream)
    {
        const int Wp = W + 2;
        const int dcv = 128 % CV, dr = 128 / CV;
        int cv = tid % CV, r = tid / CV;
        int yy = r / Wp, xx = r - yy * Wp;
        for (int i = tid; i < tile_vecs; i += 128) {
            const int gy = y0 + yy - 1, gx = xx - 1;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
            const bf16* src = in ? x + (((long long)n * H + gy) * W + gx) * C + cv * 8 : x;
            cp_async16(st + ((size_t)yy * Wp + xx) * CP + cv * 8, src, in ? 16 : 0);  // src-size 0: zero fill
            cv += dcv;
            xx += dr;
            if (cv >= CV) {
                cv -= CV;
                ++xx;
            }
            while (xx >= Wp) {
                xx -= Wp;
                ++yy;
            }
        }
        const int KV = 9 * C / 8;
        const int dkv = 128 % KV, dco = 128 / KV;
        int co = tid / KV, kv = tid - co * KV;
        for (int i = tid; i < 8 * KV; i += 128) {
            cp_async16(sw + (size_t)co * KP + kv * 8, wp + (size_t)co * 9 * C + kv * 8, 16);
            kv += dkv;
            co += dco;
            if (kv >= KV) {
                kv -= KV;
                ++co;
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- this warp's 32 pixels = two m16 tiles; each m16 tile lies inside ONE image row (16 | W): tile m = row ty[m], columns tx[m] .. + 15
    int ty[2], tx[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int pm = warp * 32 + m * 16;
        ty[m] = pm / W;
        tx[m] = pm - ty[m] * W;
    }
    float acc[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
    const uint32_t st_addr = static_cast<uint32_t>(__cvta_generic_to_shared(st));
    const uint32_t sw_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sw));
    // A fragment rows: lane & 15 = pixel within the m16 tile, (lane >> 4) * 8 = k offset; B: row = lane & 7 (co), (lane >> 3) & 1 = k half
    const int a_row = lane & 15, a_k = (lane >> 4) * 8;
    const int b_row = lane & 7, b_k = ((lane >> 3) & 1) * 8;
    const int cblocks = C / 16;
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t a_base0 = st_addr + (uint32_t)((((ty[0] + ky) * (W + 2) + tx[0] + kx + a_row) * CP + a_k) * 2);
        const uint32_t a_base1 = st_addr + (uint32_t)((((ty[1] + ky) * (W + 2) + tx[1] + kx + a_row) * CP + a_k) * 2);
        const uint32_t b_base = sw_addr + (uint32_t)((b_row * KP + tap * C + b_k) * 2);
#pragma unroll 4
        for (int cb = 0; cb < cblocks; ++cb) {
            uint32_t a0[4], a1[4], b[2];
            ldmatrix_x4(a0, a_base0 + cb * 32);
            ldmatrix_x4(a1, a_base1 + cb * 32);
            ldmatrix_x2(b, b_base + cb * 32);
            mma_bf16_16816(acc[0], a0, b);
            mma_bf16_16816(acc[1], a1, b);
        }
    }
    // ---- D fragment: rows lane / 4 (+ 8), columns (lane % 4) * 2 + {0, 1} -> out[n][co][y][x]
    const int co0 = (lane & 3) * 2;
    const long long HW = (long long)H * W;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int gy = y0 + ty[m];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int gx = tx[m] + (lane >> 2) + half * 8;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int co = co0 + j;
                if (co < Cout) out[((long long)n * Cout + co) * HW + (long long)gy * W + gx] = acc[m][half * 2 + j] + bias8[co];
            }
        }
    }
}

bool conv3x3_last_supported(int H, int W, int C, int Cout) {
    if (Cout > 8 || C % 16 || W > 128 || W % 16 || 128 % W || H % (128 / W)) return false;
    const int TH = 128 / W;
    const size_t smem = ((size_t)(TH + 2) * (W + 2) * (C + 8) + (size_t)8 * (9 * C + 8)) * sizeof(bf16);
    return smem <= 200 * 1024;
}

void conv3x3_last(const bf16* x, const bf16* wp, const float* bias8, float* out, int N, int H, int W, int C, int Cout, cudaStream_t st) {
    const int TH = 128 / W;
    const size_t smem = ((size_t)(TH + 2) * (W + 2) * (C + 8) + (size_t)8 * (9 * C + 8)) * sizeof(bf16);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(conv3x3_last_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured.set();
    }
    conv3x3_last_k<<<N * (H / TH), 128, smem, st>>>(x, wp, bias8, out, H, W, C, Cout, TH);
}

}  // namespace dxmi
