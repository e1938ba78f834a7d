// Weight gradient of a 3x3 (pad 1, stride 1) or 1x1 convolution on the tensor cores (the backward half
// of the training step, trainer.py:252-264 / :320-326):
//     dW[co][tap][ci] = sum over pixels p of dY[p, co] * X[p + tap, ci]
// Both operands are NHWC bf16, i.e. the reduction dimension (pixels) is the SLOW index of both: the tiles TMA brings in
// ([64 pixels][64 channels], SWIZZLE_128B) are "MN-major" UMMA operands - the instruction descriptor's a_major / b_major bits
// are set and the smem descriptors carry LBO = distance between 64-channel blocks, SBO = distance between 8-pixel groups.
// The filter tap is, as in the forward kernel, only a shift of the rank-4 TMA box over X (out-of-bounds rows and columns
// arrive as zeros = the padding).
// Work split: grid = (Cout/128) x taps x S pixel ranges (split-K); each CTA accumulates one 128 x Cin fp32 tile in TMEM over
// its pixel range and writes it to partial[S][Cout][taps][Cin]; wgrad_reduce_k sums the S partials in a fixed order
// (deterministic) into the OIHW fp32 gradient.
#include <cuda.h>

#include <cstdio>

#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "wgrad_tc.cuh"

namespace dxmi {

static constexpr int WG_KPIX = 64;            // pixels per K step
static constexpr int WG_BOX_BYTES = 64 * 128; // one [64 pixel][64 channel] box
static constexpr int WG_THREADS = 192;
static constexpr int WG_STAGES = 4;

struct WgradParams {
    CUtensorMap dy_map, x_map;
    float* partial;
    int Cout, Cin, taps;
    int bw, bh, bn, tiles_w, tiles_h;   // 64-pixel chunk geometry
    int nchunks, chunks_per_split, S;
};

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;   // LBO: next 64-element block along M / N
    d |= static_cast<uint64_t>(1024 >> 4) << 32;        // SBO: next group of 8 K rows (8 x 128 B)
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;                // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int a_bytes = 2 * WG_BOX_BYTES;                  // 128 output channels
    const int b_bytes = (p.Cin / 64) * WG_BOX_BYTES;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* done_bar = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int item = blockIdx.x;
    const int split = item % p.S;
    item /= p.S;
    const int tap = item % p.taps;
    const int co0 = (item / p.taps) * 128;
    const int g0 = split * p.chunks_per_split;
    int g1 = g0 + p.chunks_per_split;
    if (g1 > p.nchunks) g1 = p.nchunks;
    const int nk = g1 > g0 ? g1 - g0 : 0;
    const uint32_t tmem_cols = p.Cin <= 128 ? 128 : 256;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&p.dy_map);
        ptx::prefetch_tmap(&p.x_map);
        for (int s = 0; s < WG_STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(done_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, tmem_cols);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (ptx::elect_one()) {
            const int r = p.taps == 9 ? tap / 3 - 1 : 0, q = p.taps == 9 ? tap % 3 - 1 : 0;
            const int tpn = p.tiles_w * p.tiles_h;
            for (int k = 0; k < nk; ++k) {
                const int stage = k % WG_STAGES;
                ptx::mbar_wait(&empty_bar[stage], ((k / WG_STAGES) & 1) ^ 1);
                const int g = g0 + k;
                const int n_blk = g / tpn, rem = g - n_blk * tpn;
                const int h0 = (rem / p.tiles_w) * p.bh, w0 = (rem % p.tiles_w) * p.bw, n0 = n_blk * p.bn;
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                ptx::mbar_expect_tx(&full_bar[stage], stage_bytes);
                ptx::tma_load_4d(sa, &p.dy_map, &full_bar[stage], co0, w0, h0, n0);
                ptx::tma_load_4d(sa + WG_BOX_BYTES, &p.dy_map, &full_bar[stage], co0 + 64, w0, h0, n0);
                for (int c = 0; c < p.Cin / 64; ++c)
                    ptx::tma_load_4d(sb + c * WG_BOX_BYTES, &p.x_map, &full_bar[stage], c * 64, w0 + q, h0 + r, n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (ptx::elect_one()) {
            // kind::f16, bf16 x bf16 -> fp32, A and B both MN-major (bits 15 / 16)
            const uint32_t idesc = ptx::make_idesc(1, 128, (uint32_t)p.Cin) | (1u << 15) | (1u << 16);
            for (int k = 0; k < nk; ++k) {
                const int stage = k % WG_STAGES;
                ptx::mbar_wait(&full_bar[stage], (k / WG_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
                const uint32_t sb = sa + a_bytes;
#pragma unroll
                for (int j = 0; j < WG_KPIX / 16; ++j) {
                    const uint64_t da = make_mnmajor_sw128_desc(sa + j * 2048, WG_BOX_BYTES);
                    const uint64_t db = make_mnmajor_sw128_desc(sb + j * 2048, WG_BOX_BYTES);
                    ptx::umma_f16(tmem, da, db, idesc, (k > 0 || j > 0) ? 1u : 0u);
                }
                ptx::umma_commit(&empty_bar[stage]);
            }
            ptx::umma_commit(done_bar);
        }
        __syncwarp();
    } else {
        // epilogue: TMEM lane = output channel co0 + lane, column = input channel
        ptx::mbar_wait(done_bar, 0);
        ptx::tc_fence_after();
        const int quarter = warp & 3;
        const int co = co0 + quarter * 32 + lane;
        const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        float* dst = p.partial + (((long long)split * p.Cout + co) * p.taps + tap) * p.Cin;
        for (int c = 0; c < p.Cin; c += 32) {
            uint32_t v[32];
            if (nk > 0) {
                ptx::tmem_ld_32x32b_x32(taddr + c, v);
                ptx::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (co < p.Cout) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(dst + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, tmem_cols);
    }
}

// grad[co][ci][tap] (OIHW slice: input channels ci_off .. ci_off+Cin of a tensor with Cin_total) = sum_s partial[s][co][tap][ci]
__global__ void wgrad_reduce_k(const float* __restrict__ partial, float* __restrict__ grad, int S, int Cout, int taps, int Cin,
                               int Cin_total, int ci_off, float scale) {
    const long long total = (long long)Cout * taps * Cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int tap = (int)((i / Cin) % taps);
        const int co = (int)(i / ((long long)Cin * taps));
        float a = 0.f;
        for (int s = 0; s < S; ++s) a += partial[(long long)s * total + i];
        grad[((long long)co * Cin_total + ci_off + ci) * taps + tap] = a * scale;
    }
}

void gemm_set_error(const char* msg);

int prepare_wgrad(const void* dy, const void* x, int N, int H, int W, int Cout, int Cin, int taps, WgradOp* op, long long dy_ld,
                  long long x_ld) {
    if (dy_ld <= 0) dy_ld = Cout;
    if (x_ld <= 0) x_ld = Cin;
    if (Cout % 128 || (Cin != 64 && Cin != 128 && Cin != 192 && Cin != 256) || (taps != 1 && taps != 9)) {
        gemm_set_error("wgrad: need Cout % 128 == 0, Cin in {64,128,192,256}, taps 1 or 9");
        return -30;
    }
    const int bw = W < 64 ? W : 64;
    int bh = 64 / bw;
    if (bh > H) bh = H;
    const int bn = 64 / (bw * bh);
    if (bw * bh * bn != 64 || W % bw || H % bh) {
        gemm_set_error("wgrad: unsupported map geometry (need power-of-two 64-pixel chunks)");
        return -31;
    }
    WgradParams& p = *reinterpret_cast<WgradParams*>(op->params);
    static_assert(sizeof(WgradParams) <= sizeof(op->params), "WgradOp::params too small");
    int r = make_act_map(&p.dy_map, dy, Cout, W, H, N, dy_ld, dy_ld * W, dy_ld * W * H, bw, bh, bn, 1);
    if (r) return r;
    r = make_act_map(&p.x_map, x, Cin, W, H, N, x_ld, x_ld * W, x_ld * W * H, bw, bh, bn, 1);
    if (r) return r;
    p.Cout = Cout;
    p.Cin = Cin;
    p.taps = taps;
    p.bw = bw;
    p.bh = bh;
    p.bn = bn;
    p.tiles_w = W / bw;
    p.tiles_h = H / bh;
    p.nchunks = ((N + bn - 1) / bn) * p.tiles_w * p.tiles_h;
    // split-K: one wave of CTAs (the smem ring allows one CTA per SM and a CTA's prologue / TMEM drain are not overlapped
    // with anything, so a second partial wave only adds fixed cost), at least 4 K steps each
    const int base_items = (Cout / 128) * taps;
    int S = 148 / base_items;
    if (S > p.nchunks / 4) S = p.nchunks / 4;
    if (S < 1) S = 1;
    p.chunks_per_split = (p.nchunks + S - 1) / S;
    S = (p.nchunks + p.chunks_per_split - 1) / p.chunks_per_split;
    p.S = S;
    p.partial = nullptr;
    op->S = S;
    op->Cout = Cout;
    op->Cin = Cin;
    op->taps = taps;
    op->grid = base_items * S;
    op->smem = WG_STAGES * (2 * WG_BOX_BYTES + (Cin / 64) * WG_BOX_BYTES) + 256;
    op->partial_floats = (size_t)S * Cout * taps * Cin;
    op->flops = 2.0 * N * H * W * (double)Cout * Cin * taps;
    return 0;
}

int run_wgrad(const WgradOp& op, float* partial_ws, float* grad, int Cin_total, int ci_off, float scale, cudaStream_t st) {
    static DevFlags configured;
    if (!configured.test()) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            gemm_set_error(cudaGetErrorString(e));
            return (int)e;
        }
        configured.set();
    }
    WgradParams p = *reinterpret_cast<const WgradParams*>(op.params);
    p.partial = partial_ws;
    wgrad_kernel<<<op.grid, WG_THREADS, op.smem, st>>>(p);
    const long long total = (long long)op.Cout * op.taps * op.Cin;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    wgrad_reduce_k<<<blocks, 256, 0, st>>>(partial_ws, grad, op.S, op.Cout, op.taps, op.Cin, Cin_total, ci_off, scale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gemm_set_error(cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace dxmi
