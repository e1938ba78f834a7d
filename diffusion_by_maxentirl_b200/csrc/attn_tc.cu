// Fused multi-head self-attention (head dim 64) on tcgen05 tensor cores - QKVAttentionLegacy.forward
// (models/cm/unet.py:413-441) for sequence lengths that are multiples of 128 (16x16 and 32x32 feature maps) and for
// seq == 64 (8x8 maps: one 64-key tile; the upper 64 query rows of the 128-row MMA tile are zero-filled padding).
//
//   grid = (seq / 128, heads, batch); one CTA owns 128 query rows of one (image, head) and streams the keys in
//   tiles of 128:   S = Q K^T  (tcgen05, M=128 N=128 K=64, fp32 in TMEM)
//                   P = exp2(S * scale*log2e - m)   (4 softmax warps: thread <-> query row <-> TMEM lane)
//                   O_tile = P V  (tcgen05, M=128 N=64 K=128; P staged bf16 in smem in the SWIZZLE_128B K-major
//                   layout a TMA load would have produced; V^T comes from the [C, seq] "value-transposed" GEMM)
//   and the running (m, l, O) online-softmax state lives in registers (O_tile is read back from TMEM per key tile,
//   so no TMEM read-modify-write is needed).
//
// Warp roles (192 threads): warps 0-3 softmax + output, warp 4 TMA producer, warp 5 TMEM allocator + MMA issuer.
// K/V tiles are double buffered; S_{j+1} is issued right behind P V_j so the tensor pipe works while the softmax
// warps fold O_j into their registers.
// (Round 2, measured negative: 8 softmax warps - two per TMEM lane quarter, each half of the key columns, row max exchanged
//  through shared memory - ran seq 1024 at 343 us instead of 239 us: the per-tile chain softmax -> P V -> fold is bound by its
//  synchronisation latencies, not by issue slots; the fix is S double-buffering with two query tiles in flight, not more warps.)
#include "attn_tc.cuh"
#include "ptx.cuh"

#include <cstdio>

namespace dxmi {

static constexpr int ATT_THREADS = 192;
static constexpr int ATT_D = 64;
static constexpr int ATT_TILE = 128;
static constexpr int SM_Q = 0;
static constexpr int SM_K = 16 * 1024;            // 2 stages x 16 KB  [128 keys x 64 d]
static constexpr int SM_V = SM_K + 2 * 16 * 1024; // 2 stages x 16 KB  2 x [64 d x 64 keys]
static constexpr int SM_P = SM_V + 2 * 16 * 1024; // 32 KB             2 x [128 rows x 64 keys]
static constexpr int SM_BAR = SM_P + 32 * 1024; // 8 mbarriers + TMEM slot
static constexpr int ATT_SMEM = SM_BAR + 128;     // 112 KB + 128 B
static constexpr uint32_t TM_S = 0, TM_O = 128, TM_COLS = 256;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // barriers live behind the tiles inside the dynamic allocation (no static smem: two CTAs must fit one SM)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t& q_full = bars[0];
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 3;
    uint64_t& s_full = bars[5];
    uint64_t& p_full = bars[6];
    uint64_t& o_full = bars[7];
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int kt = p.seq < ATT_TILE ? p.seq : ATT_TILE;  // keys per tile: 64 or 128
    const int n_tiles = p.seq / kt;

    if (threadIdx.x == 0) {
        if (ptx::smem_u32(smem) & 1023u) {
            printf("dxmi attn: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tmap(&p.qk_map);
        ptx::prefetch_tmap(&p.vt_map);
        ptx::mbar_init(&q_full, 1);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&kv_full[s], 1);
            ptx::mbar_init(&kv_empty[s], 1);
        }
        ptx::mbar_init(&s_full, 1);
        ptx::mbar_init(&p_full, 128);
        ptx::mbar_init(&o_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 5) ptx::tmem_alloc(&tmem_slot, TM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&q_full, 16 * 1024);
            ptx::tma_load_3d(smem + SM_Q, &p.qk_map, &q_full, p.q_col0 + head * ATT_D, q_tile * ATT_TILE, b);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                ptx::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                if (p.v_col0 >= 0) {
                    // V straight from the fused q|k|v projection: [keys][64 d] rows of 128 bytes (MN-major B operand of P.V)
                    ptx::mbar_expect_tx(&kv_full[s], 32 * 1024);
                    ptx::tma_load_3d(smem + SM_K + s * 16384, &p.qk_map, &kv_full[s], p.k_col0 + head * ATT_D, j * kt, b);
                    ptx::tma_load_3d(smem + SM_V + s * 16384, &p.qk_map, &kv_full[s], p.v_col0 + head * ATT_D, j * kt, b);
                    continue;
                }
                ptx::mbar_expect_tx(&kv_full[s], kt == ATT_TILE ? 32 * 1024 : 24 * 1024);
                ptx::tma_load_3d(smem + SM_K + s * 16384, &p.qk_map, &kv_full[s], p.k_col0 + head * ATT_D, j * kt, b);
                ptx::tma_load_3d(smem + SM_V + s * 16384, &p.vt_map, &kv_full[s], j * kt, head * ATT_D, b);
                if (kt == ATT_TILE)
                    ptx::tma_load_3d(smem + SM_V + s * 16384 + 8192, &p.vt_map, &kv_full[s], j * kt + 64, head * ATT_D, b);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer
        if (ptx::elect_one()) {
            const uint32_t idesc_s = ptx::make_idesc(1, 128, kt);
            constexpr uint32_t idesc_o = ptx::make_idesc(1, 128, 64);
            const int pv_steps = kt / 16;
            const uint64_t dq = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + SM_Q));
            const uint32_t sp = ptx::smem_u32(smem + SM_P);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                ptx::mbar_wait(&kv_full[s], (j >> 1) & 1);
                ptx::tc_fence_after();
                const uint64_t dk = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + SM_K + s * 16384));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
                ptx::umma_commit(&s_full);
            };
            ptx::mbar_wait(&q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                ptx::mbar_wait(&p_full, j & 1);
                ptx::tc_fence_after();
                const uint32_t sv = ptx::smem_u32(smem + SM_V + s * 16384);
                if (p.v_col0 >= 0) {
                    // B = V tile [kt keys][64 d]: the reduction index (keys) is the ROW index -> MN-major operand (b_major bit 16);
                    // 16 keys per MMA = 2 KB of rows; SBO = 1 KB between 8-key groups (same encoding as wgrad_tc.cu)
                    for (int k = 0; k < pv_steps; ++k) {
                        const uint64_t da = ptx::make_kmajor_sw128_desc(sp + (k >> 2) * 16384) + 2 * (k & 3);
                        uint64_t db = 0;
                        const uint32_t addr = sv + k * 2048;
                        db |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
                        db |= static_cast<uint64_t>(16384 >> 4) << 16;  // LBO (one 64-wide block only)
                        db |= static_cast<uint64_t>(1024 >> 4) << 32;   // SBO
                        db |= static_cast<uint64_t>(1) << 46;
                        db |= static_cast<uint64_t>(2) << 61;           // SWIZZLE_128B
                        ptx::umma_f16(tmem + TM_O, da, db, idesc_o | (1u << 16), k > 0);
                    }
                } else {
                    for (int k = 0; k < pv_steps; ++k) {
                        const uint64_t da = ptx::make_kmajor_sw128_desc(sp + (k >> 2) * 16384) + 2 * (k & 3);
                        const uint64_t db = ptx::make_kmajor_sw128_desc(sv + (k >> 2) * 8192) + 2 * (k & 3);
                        ptx::umma_f16(tmem + TM_O, da, db, idesc_o, k > 0);
                    }
                }
                ptx::umma_commit(&o_full);
                ptx::umma_commit(&kv_empty[s]);
                if (j + 1 < n_tiles) issue_s(j + 1);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax + output (thread <-> query row)
        const int row = warp * 32 + lane;
        const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        uint8_t* prow = smem + SM_P + (row >> 3) * 1024 + (row & 7) * 128;
        const int sw = row & 7;
        float m = -INFINITY, l = 0.f;
        float o[ATT_D];
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;

        for (int j = 0; j < n_tiles; ++j) {
            ptx::mbar_wait(&s_full, j & 1);
            ptx::tc_fence_after();
            // pass 1: row max of this key tile
            const int ncol32 = kt >> 5;
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < ncol32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m, mx * p.scale_log2);
            const float alpha = fast_exp2(m - m_new);
            // pass 2: probabilities -> bf16 -> swizzled smem (A operand of P V)
            float rs = 0.f;
#pragma unroll 1
            for (int c = 0; c < ncol32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_new));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_new));
                    rs += p0 + p1;
                    __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
                    pk[i] = *reinterpret_cast<uint32_t*>(&t);
                }
                // keys c*32 .. c*32+31 -> 64-key chunk (c >> 1), 16-byte units ((c & 1) * 4 + u), XOR-swizzled by row
                uint8_t* dst = prow + (c >> 1) * 16384;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int unit = ((c & 1) * 4 + u) ^ sw;
                    *reinterpret_cast<uint4*>(dst + unit * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
            l = l * alpha + rs;
            m = m_new;
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&p_full);
            // fold O_j into the running output
            ptx::mbar_wait(&o_full, j & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_row + TM_O + c * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(v[i]));
            }
        }
        const float inv = 1.f / l;
        __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.seq + q_tile * ATT_TILE + row) * p.ldo + head * ATT_D;
        if (q_tile * ATT_TILE + row < p.seq) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(o[q * 8 + 0] * inv, o[q * 8 + 1] * inv);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(o[q * 8 + 2] * inv, o[q * 8 + 3] * inv);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(o[q * 8 + 4] * inv, o[q * 8 + 5] * inv);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(o[q * 8 + 6] * inv, o[q * 8 + 7] * inv);
            u.x = *reinterpret_cast<uint32_t*>(&t0);
            u.y = *reinterpret_cast<uint32_t*>(&t1);
            u.z = *reinterpret_cast<uint32_t*>(&t2);
            u.w = *reinterpret_cast<uint32_t*>(&t3);
            reinterpret_cast<uint4*>(dst)[q] = u;
        }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, TM_COLS);
    }
}

static thread_local char g_attn_err[384] = "";
const char* attn_last_error() { return g_attn_err; }

int prepare_attn(const void* qk, long long ld_qk, int q_col0, int k_col0, const void* vt, void* out, int ldo, int B,
                 int heads, int seq, int d, float scale, AttnOp* op, int v_col0) {
    g_attn_err[0] = 0;
    if (d != ATT_D || (seq % ATT_TILE && seq != 64) || (ldo & 7)) {
        snprintf(g_attn_err, sizeof g_attn_err, "fused attention needs head dim 64 and seq == 64 or seq %% 128 == 0 (got d=%d seq=%d)", d, seq);
        return -30;
    }
    AttnParams& p = op->p;
    const int C = heads * d;
    // q | k block: [B, seq, ld_qk] bf16; box = 64 columns x 128 rows
    int r = make_mat_map(&p.qk_map, qk, (int)ld_qk, seq, B, ld_qk, (long long)seq * ld_qk, ATT_TILE);
    if (r) {
        snprintf(g_attn_err, sizeof g_attn_err, "%s", gemm_last_error());
        return r;
    }
    if (vt) {
        // V^T: [B, C, seq] bf16; box = 64 keys x 64 channel rows
        r = make_mat_map(&p.vt_map, vt, seq, C, B, seq, (long long)C * seq, ATT_D);
        if (r) {
            snprintf(g_attn_err, sizeof g_attn_err, "%s", gemm_last_error());
            return r;
        }
        v_col0 = -1;
    } else {
        if (v_col0 < 0) {
            snprintf(g_attn_err, sizeof g_attn_err, "prepare_attn: neither V^T nor a V column offset given");
            return -31;
        }
        p.vt_map = p.qk_map;
    }
    p.v_col0 = v_col0;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.seq = seq;
    p.q_col0 = q_col0;
    p.k_col0 = k_col0;
    p.scale_log2 = scale * 1.4426950408889634f;
    op->grid = dim3((seq + ATT_TILE - 1) / ATT_TILE, heads, B);
    op->flops = 4.0 * B * heads * (double)seq * seq * d;
    return 0;
}

int run_attn(const AttnOp& op, cudaStream_t st) {
    static DevFlags configured;
    if (!configured.test()) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
        if (e != cudaSuccess) {
            snprintf(g_attn_err, sizeof g_attn_err, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured.set();
    }
    attn_fwd_kernel<<<op.grid, ATT_THREADS, ATT_SMEM, st>>>(op.p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_attn_err, sizeof g_attn_err, "attn launch: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace dxmi
