// Last convolution of a U-Net: 3x3, stride 1, pad 1, C -> Cout <= 8 channels (3 in every DxMI config), NHWC bf16 in, fp32 NCHW out
// (unet_small.py:330-332 conv_out, cm/unet.py:738-742 `out`).
//
// Why not the tcgen05 GEMM: with 3 output channels the contraction is 0.03 % of a U-Net's FLOPs but the implicit-GEMM kernels
// pay for it like a 32-column tile AND re-read the input once per tap through L2 (9 x 67 MB at CIFAR B = 256: 74-92 us).  The
// operation is a read-once HBM stream: this kernel stages a (rows + 2) x (W + 2) x C halo tile of the input in shared memory ONCE (TMA),
// keeps the 8 x 9C weights next to it, and contracts with warp-level mma.sync.m16n8k16 - the one place where the legacy tensor-core
// instruction is the right tool: its N = 8 tile is exactly the (zero-padded) output width, there is no accumulator to drain and
// no epilogue worth overlapping.  One CTA = 128 output pixels (TH full image rows), one warp = 32 pixels = two m16 tiles (each inside one image row: 16 | W).
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

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

// xmap: (c, w, h, n) map over x [N,H,W,C] bf16 with box (64, W + 2, TH + 2, 1), SWIZZLE_128B; wp [8][9*C] bf16 (row co, k = tap*C + c,
// rows >= Cout zero); bias8 [8] fp32; out [N,Cout,H,W] fp32.  grid = N * H / TH, block = 128, TH * W == 128.
// Staging: ONE elected thread issues C/64 TMA box loads - the halo tile [(TH+2) x (W+2) pixels] of each 64-channel slice, zero-filled
// outside the image (= the padding) - into [slice][pixel][128 B] rows with the 128-byte XOR swizzle, which makes the ldmatrix reads
// below conflict-free without padding.  (History, ncu on CIFAR B = 256, a 67 MB read-once stream: register-staged loads 102 us;
// cp.async 58 us, still issue bound - 3.4 k instructions per warp, most of them address arithmetic of the copy loop.)
__global__ void __launch_bounds__(128) conv3x3_last_k(const __grid_constant__ CUtensorMap xmap, const bf16* __restrict__ wp,
                                                      const float* __restrict__ bias8, float* __restrict__ out, int H, int W, int C, int Cout,
                                                      int TH, int slice_stride) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int KP = 9 * C + 8;        // padded weight-row pitch (elements): conflict-free ldmatrix rows
    const int nsl = C >> 6;
    bf16* sw = reinterpret_cast<bf16*>(sm + (size_t)nsl * slice_stride);   // [8][KP]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (size_t)nsl * slice_stride + (size_t)8 * KP * sizeof(bf16));
    const int tiles_per_img = H / TH;
    const int n = blockIdx.x / tiles_per_img;
    const int y0 = (blockIdx.x - n * tiles_per_img) * TH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Wp = W + 2;
    if (tid == 0) {
        if (ptx::smem_u32(sm) & 1023u) __trap();
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
        ptx::mbar_expect_tx(bar, (uint32_t)nsl * (TH + 2) * Wp * 128);
        for (int sl = 0; sl < nsl; ++sl) ptx::tma_load_4d(sm + (size_t)sl * slice_stride, &xmap, bar, sl * 64, -1, y0 - 1, n);
    }
    {
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
    __syncthreads();  // weights staged; the barrier is initialised
    ptx::mbar_wait(bar, 0);
    // ---- this warp's 32 pixels = two m16 tiles; each m16 tile lies inside ONE image row (16 | W): tile m = row ty[m], columns tx[m] .. + 15
    int ty[2], tx[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int pm = warp * 32 + m * 16;
        ty[m] = pm / W;
        tx[m] = pm - ty[m] * W;
    }
    float acc[2][3][4];  // [m tile][kernel row: three independent accumulation chains][fragment]
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[m][r][j] = 0.f;
    const uint32_t st_addr = ptx::smem_u32(sm);
    const uint32_t sw_addr = ptx::smem_u32(sw);
    // A fragment rows: lane & 15 = pixel within the m16 tile, lane >> 4 = which 8-channel unit of the k16 step; B: row = lane & 7 (co), (lane >> 3) & 1 = k half
    const int a_row = lane & 15, a_hi = lane >> 4;
    const int b_row = lane & 7, b_k = ((lane >> 3) & 1) * 8;
    const int cblocks = C / 16;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
        for (int kx = 0; kx < 3; ++kx) {
            const int p0 = (ty[0] + ky) * Wp + tx[0] + kx + a_row, p1 = (ty[1] + ky) * Wp + tx[1] + kx + a_row;
            const uint32_t row0 = st_addr + p0 * 128, row1 = st_addr + p1 * 128;
            const int sw0 = p0 & 7, sw1 = p1 & 7;
            const uint32_t b_base = sw_addr + (uint32_t)((b_row * KP + (ky * 3 + kx) * C + b_k) * 2);
#pragma unroll 4
            for (int cb = 0; cb < cblocks; ++cb) {
                const uint32_t sl_off = (uint32_t)(cb >> 2) * slice_stride;
                const int unit = (cb & 3) * 2 + a_hi;
                uint32_t a0[4], a1[4], b[2];
                ldmatrix_x4(a0, row0 + sl_off + ((unit ^ sw0) << 4));
                ldmatrix_x4(a1, row1 + sl_off + ((unit ^ sw1) << 4));
                ldmatrix_x2(b, b_base + cb * 32);
                mma_bf16_16816(acc[0][ky], a0, b);
                mma_bf16_16816(acc[1][ky], a1, b);
            }
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
                const float v = (acc[m][0][half * 2 + j] + acc[m][1][half * 2 + j]) + acc[m][2][half * 2 + j];
                if (co < Cout) out[((long long)n * Cout + co) * HW + (long long)gy * W + gx] = v + bias8[co];
            }
        }
    }
}

static size_t conv_last_slice_stride(int W, int TH) { return ((size_t)(TH + 2) * (W + 2) * 128 + 1023) & ~(size_t)1023; }
static size_t conv_last_smem(int W, int C, int TH) { return (size_t)(C / 64) * conv_last_slice_stride(W, TH) + (size_t)8 * (9 * C + 8) * sizeof(bf16) + 16; }

bool conv3x3_last_supported(int H, int W, int C, int Cout) {
    if (Cout > 8 || C % 64 || W > 128 || W % 16 || 128 % W || H % (128 / W)) return false;
    return conv_last_smem(W, C, 128 / W) <= 200 * 1024;
}

int prepare_conv3x3_last(const bf16* x, int N, int H, int W, int C, ConvLastOp* op) {
    op->N = N;
    op->H = H;
    op->W = W;
    op->C = C;
    const int TH = 128 / W;
    return make_act_map(&op->xmap, x, C, W, H, N, C, (long long)W * C, (long long)H * W * C, W + 2, TH + 2, 1, 1);
}

void conv3x3_last(const ConvLastOp& op, const bf16* wp, const float* bias8, float* out, int Cout, cudaStream_t st) {
    const int TH = 128 / op.W;
    const size_t smem = conv_last_smem(op.W, op.C, TH);
    static DevFlags configured;
    if (!configured.test()) {
        cudaFuncSetAttribute(conv3x3_last_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured.set();
    }
    conv3x3_last_k<<<op.N * (op.H / TH), 128, smem, st>>>(op.xmap, wp, bias8, out, op.H, op.W, op.C, Cout, TH, (int)conv_last_slice_stride(op.W, TH));
}

}  // namespace dxmi
