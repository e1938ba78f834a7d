// Fused multi-head self-attention (head dim 64) on tcgen05 tensor cores - QKVAttentionLegacy.forward
// (models/cm/unet.py:413-441) for sequence lengths that are multiples of 128 (16x16 and 32x32 feature maps) and for
// seq == 64 (8x8 maps: one 64-key tile; the upper 64 query rows of the 128-row MMA tile are zero-filled padding).
//
//   grid = (seq / 128, heads, batch); one CTA owns 128 query rows of one (image, head) and streams the keys in
//   tiles of 128:   S = Q K^T  (tcgen05, M=128 N=128 K=64, fp32 in TMEM)
//                   P = exp2(S * scale*log2e - m)   (4 softmax warps: thread <-> query row <-> TMEM lane)
//                   O += P V  (tcgen05, M=128 N=64 K=128, accumulated IN TMEM over all key tiles; P is written back to tensor memory as
//                   bf16 pairs and enters the MMA as a TMEM A operand - no shared-memory staging, no proxy fence: 181 -> 171 us at seq 1024)
//
// Warp roles (192 threads): warps 0-3 softmax + output, warp 4 TMA producer, warp 5 TMEM allocator + MMA issuer.
// Round-2 rework of the per-tile chain (the first version ran seq 1024 at 0.43 PFLOP/s: its softmax threads read S twice from TMEM
// - a max pass and an exp pass -, waited for P V, read O back and folded it into 64 registers, every tile, all on one serial chain):
//   * the whole S row (128 fp32) is loaded ONCE into registers - O lives in TMEM now, which frees the registers for it - by four
//     asynchronous tcgen05.ld and one wait; the thread then releases the S buffer (barrier s_free) BEFORE it computes anything, so
//     the MMA warp issues S_{j+1} = Q K_{j+1}^T while the exponentials of tile j are still being evaluated;
//   * lazy rescaling (as in FlashAttention-4): the reference point m of the exponent only moves when the row maximum grew by
//     more than 2^8 since it was last set (warp-uniform decision), so O (in TMEM) and l are rescaled - a tcgen05.ld / .st round
//     trip - a handful of times per row instead of once per tile; P <= 256 keeps full bf16 relative precision, the final
//     O / l is exact;
//   * P V_j only has to be finished before P_{j+1} is WRITTEN (one tile later), so it overlaps the next tile's load + max.
// K/V tiles are double buffered.
// (Round 2, measured negative on the first version: 8 softmax warps - two per TMEM lane quarter - ran seq 1024 at 343 us instead of
//  239 us: the chain was bound by its synchronisation latencies, not by issue slots.)
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
static constexpr int SM_BAR = SM_V + 2 * 16 * 1024; // 13 mbarriers + TMEM slot
static constexpr int ATT_SMEM = SM_BAR + 128;     // 80 KB + 128 B (P lives in tensor memory)
static constexpr uint32_t TM_S = 0, TM_O = 128, TM_P = 192, TM_COLS = 256;  // P: 128 keys as 64 columns of bf16 pairs (A operand of P V, read from TMEM)

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// KT = keys per tile: 128, or 64 for seq == 64 (compile time: with a run-time tile width the unrolled softmax loops degenerate into
// one branch + one exposed MUFU latency per pair of exponentials - 4800 instead of ~1100 cycles per tile, measured with clock64)
// 2^x on the FMA / ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial, rel. error 2.4e-4 - far below the bf16
// rounding of P): the MUFU unit evaluates 16 ex2 per clock and SM, which bounds a d = 64 attention tile at 1024 cycles against
// 512 cycles of tensor work, so a fraction of the exponentials is computed here instead (FlashAttention-4's trick).
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part n = round(x) lands in the low mantissa bits
    const float f = x - (t - 12582912.f);  // in [-0.5, 0.5]
    const float p = fmaf(fmaf(fmaf(0.052731406f, f, 0.24209385f), f, 0.69359708f), f, 0.99996608f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // p * 2^n
}
#ifndef ATT_POLY_MASK
#define ATT_POLY_MASK (-1)  // 3: pairs whose index has (i & 3) == 3 take the polynomial (a quarter of the exponentials).  MEASURED on B200:
// 185.9 us instead of 180.8 us at seq 1024 - the kernel is bound by the serial per-warp chain, not by MUFU throughput; off
#endif

// (Measured neutral and removed: delaying every second CTA of an SM by half a tile so that one CTA's exponentials run under the
//  other's load / max / store phases - 183 vs 181 us.)
template <int KT>
__global__ void __launch_bounds__(ATT_THREADS, 2) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // barriers live behind the tiles inside the dynamic allocation (no static smem: two CTAs must fit one SM)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    // K and V stages are released separately: a K tile is dead as soon as S_j = Q K_j^T retired, a whole softmax earlier than its V
    // tile (after P V_j) - so K_{j+2} is requested two tiles ahead although each operand has only two stages
    uint64_t& q_full = bars[0];
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 3;
    uint64_t& s_full = bars[5];
    uint64_t& p_full = bars[6];
    uint64_t& o_full = bars[7];   // P V_j retired (its P tile and the O accumulator may be touched)
    uint64_t& s_free = bars[8];   // every softmax thread holds its S row in registers: the S accumulator may be overwritten
    uint64_t* v_full = bars + 9;
    uint64_t* v_empty = bars + 11;
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    constexpr int kt = KT;  // keys per tile: 64 or 128
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
            ptx::mbar_init(&k_full[s], 1);
            ptx::mbar_init(&k_empty[s], 1);
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&v_empty[s], 1);
        }
        ptx::mbar_init(&s_full, 1);
        ptx::mbar_init(&p_full, 128);
        ptx::mbar_init(&o_full, 1);
        ptx::mbar_init(&s_free, 128);
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
            auto load_k = [&](int j) {
                const int s = j & 1;
                ptx::mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
                ptx::mbar_expect_tx(&k_full[s], 16 * 1024);  // (the whole 128-row box counts, rows past seq = 64 are zero-filled)
                ptx::tma_load_3d(smem + SM_K + s * 16384, &p.qk_map, &k_full[s], p.k_col0 + head * ATT_D, j * kt, b);
            };
            auto load_v = [&](int j) {
                const int s = j & 1;
                ptx::mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
                ptx::mbar_expect_tx(&v_full[s], (p.v_col0 >= 0 || kt == ATT_TILE) ? 16 * 1024 : 8 * 1024);
                if (p.v_col0 >= 0) {
                    // V straight from the fused q|k|v projection: [keys][64 d] rows of 128 bytes (MN-major B operand of P.V)
                    ptx::tma_load_3d(smem + SM_V + s * 16384, &p.qk_map, &v_full[s], p.v_col0 + head * ATT_D, j * kt, b);
                } else {
                    ptx::tma_load_3d(smem + SM_V + s * 16384, &p.vt_map, &v_full[s], j * kt, head * ATT_D, b);
                    if (kt == ATT_TILE)
                        ptx::tma_load_3d(smem + SM_V + s * 16384 + 8192, &p.vt_map, &v_full[s], j * kt + 64, head * ATT_D, b);
                }
            };
            // issue order = the order in which the stages come free: K runs one tile ahead of V
            load_k(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) load_k(j + 1);
                load_v(j);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s = ptx::make_idesc(1, 128, kt);
            constexpr uint32_t idesc_o = ptx::make_idesc(1, 128, 64);
            constexpr int pv_steps = kt / 16;
            const uint64_t dq = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + SM_Q));
            auto issue_s = [&](int j) {
                const int s = j & 1;
                ptx::mbar_wait(&k_full[s], (j >> 1) & 1);
                ptx::tc_fence_after();
                const uint64_t dk = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + SM_K + s * 16384));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
                ptx::umma_commit(&s_full);
                ptx::umma_commit(&k_empty[s]);
            };
            ptx::mbar_wait(&q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                if (j + 1 < n_tiles) {
                    // S_{j+1} as soon as S_j sits in the softmax threads' registers: it runs under their exponentials
                    ptx::mbar_wait(&s_free, j & 1);
                    ptx::tc_fence_after();
                    issue_s(j + 1);
                }
                ptx::mbar_wait(&p_full, j & 1);
                ptx::mbar_wait(&v_full[s], (j >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t sv = ptx::smem_u32(smem + SM_V + s * 16384);
                if (p.v_col0 >= 0) {
                    // B = V tile [kt keys][64 d]: the reduction index (keys) is the ROW index -> MN-major operand (b_major bit 16);
                    // 16 keys per MMA = 2 KB of rows; SBO = 1 KB between 8-key groups (same encoding as wgrad_tc.cu)
                    for (int k = 0; k < pv_steps; ++k) {
                        uint64_t db = 0;
                        const uint32_t addr = sv + k * 2048;
                        db |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
                        db |= static_cast<uint64_t>(16384 >> 4) << 16;  // LBO (one 64-wide block only)
                        db |= static_cast<uint64_t>(1024 >> 4) << 32;   // SBO
                        db |= static_cast<uint64_t>(1) << 46;
                        db |= static_cast<uint64_t>(2) << 61;           // SWIZZLE_128B
                        ptx::umma_f16_ts(tmem + TM_O, tmem + TM_P + k * 8, db, idesc_o | (1u << 16), (j > 0 || k > 0) ? 1u : 0u);
                    }
                } else {
                    for (int k = 0; k < pv_steps; ++k) {
                        const uint64_t db = ptx::make_kmajor_sw128_desc(sv + (k >> 2) * 8192) + 2 * (k & 3);
                        ptx::umma_f16_ts(tmem + TM_O, tmem + TM_P + k * 8, db, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                    }
                }
                ptx::umma_commit(&o_full);
                ptx::umma_commit(&v_empty[s]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax + output (thread <-> query row)
        const int row = warp * 32 + lane;
        const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        constexpr float RESCALE_LOG2 = 8.f;  // move the exponent's reference point only when the row maximum grew by more than 2^8
        float m = -INFINITY, l = 0.f;        // m: reference point of every exponential taken so far (>= row max - 8)
#ifdef ATT_PROFILE
        long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
        int n_resc = 0;
#define ATT_T(k) { const long long now_ = clock64(); prof[k] += now_ - tprev; tprev = now_; }
#else
#define ATT_T(k)
#endif

        for (int j = 0; j < n_tiles; ++j) {
            ptx::mbar_wait(&s_full, j & 1);
            ptx::tc_fence_after();
            ATT_T(0);
            // the whole S row of this key tile: four asynchronous TMEM loads, one wait
            uint32_t v[ATT_TILE];
            {
                uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                uint32_t(&v2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[64]);
                uint32_t(&v3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[96]);
                ptx::tmem_ld_32x32b_x32(t_row + TM_S, v0);
                ptx::tmem_ld_32x32b_x32(t_row + TM_S + 32, v1);
                if (kt == ATT_TILE) {
                    ptx::tmem_ld_32x32b_x32(t_row + TM_S + 64, v2);
                    ptx::tmem_ld_32x32b_x32(t_row + TM_S + 96, v3);
                }
                ptx::tmem_ld_wait();
            }
            ATT_T(1);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&s_free);
            // (eight independent chains: one serial fmax chain over 128 registers is 500 cycles of latency)
            float mx8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(v[i]);
#pragma unroll
            for (int i = 8; i < 64; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(v[i]));
            if (kt == ATT_TILE) {
#pragma unroll
                for (int i = 64; i < ATT_TILE; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(v[i]));
            }
            const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]))) * p.scale_log2;
            ATT_T(2);
            // P V_{j-1} must have retired before this tile's P overwrites the staging buffer (and before O is rescaled)
            bool pv_waited = false;
            if (__any_sync(0xffffffffu, mx > m + RESCALE_LOG2)) {
                // (warp-uniform: tcgen05.ld / .st are warp collectives; every lane of the warp moves its reference point)
                const float m_new = fmaxf(m, mx);
                const float alpha = fast_exp2(m - m_new);  // 0 on the first tile (m = -inf)
                l *= alpha;
                m = m_new;
                if (j > 0) {
                    ptx::mbar_wait(&o_full, (j - 1) & 1);
                    ptx::tc_fence_after();
                    pv_waited = true;
                    // (8 columns at a time: the 128 S registers stay live across this rare block)
#pragma unroll 1
                    for (int c = 0; c < ATT_D / 8; ++c) {
                        uint32_t o[8];
                        ptx::tmem_ld_32x32b_x8(t_row + TM_O + c * 8, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        ptx::tmem_st_32x32b_x8(t_row + TM_O + c * 8, o);
                    }
                    ptx::tmem_st_wait();
                }
            }
            ATT_T(3);
            // probabilities -> bf16 (in place over the S registers, two per word)
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
            constexpr int nw = kt >> 1;
#pragma unroll
            for (int i = 0; i < ATT_TILE / 2; ++i) {
                if (i < nw) {
                    const float x0 = fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m), x1 = fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m);
                    const bool poly = ATT_POLY_MASK >= 0 && (i & 3) == ATT_POLY_MASK;
                    const float p0 = poly ? poly_exp2(x0) : fast_exp2(x0);
                    const float p1 = poly ? poly_exp2(x1) : fast_exp2(x1);
                    rs4[i & 3] += p0 + p1;
                    __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
                    v[i] = *reinterpret_cast<uint32_t*>(&t);
                }
            }
            l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            ATT_T(4);
            if (j > 0 && !pv_waited) {
                ptx::mbar_wait(&o_full, (j - 1) & 1);
                ptx::tc_fence_after();
            }
            ATT_T(5);
            // P -> tensor memory (columns TM_P ..: word i = keys 2i, 2i+1 of this thread's row), the A operand of P V: no shared-memory
            // round trip (256 cycles of smem write bandwidth per tile) and no proxy fence
            {
                uint32_t(&w0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                ptx::tmem_st_32x32b_x32(t_row + TM_P, w0);
                if (kt == ATT_TILE) {
                    uint32_t(&w1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                    ptx::tmem_st_32x32b_x32(t_row + TM_P + 32, w1);
                }
                ptx::tmem_st_wait();
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&p_full);
            ATT_T(6);
        }
#ifdef ATT_PROFILE
        if (threadIdx.x == 0 && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 3)
            printf("attn profile (cycles over %d tiles): wait S %lld | ld S %lld | max %lld | rescale %lld | exp %lld | wait PV %lld | store P %lld\n", n_tiles,
                   prof[0], prof[1], prof[2], prof[3], prof[4], prof[5], prof[6]);
#endif
        // O = (sum_j P_j V_j) / l
        ptx::mbar_wait(&o_full, (n_tiles - 1) & 1);
        ptx::tc_fence_after();
        const float inv = 1.f / l;
        __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.seq + q_tile * ATT_TILE + row) * p.ldo + head * ATT_D;
        const bool row_ok = q_tile * ATT_TILE + row < p.seq;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            ptx::tmem_ld_32x32b_x32(t_row + TM_O + c * 32, o);
            ptx::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
                    __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
                    __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
                    u.x = *reinterpret_cast<uint32_t*>(&t0);
                    u.y = *reinterpret_cast<uint32_t*>(&t1);
                    u.z = *reinterpret_cast<uint32_t*>(&t2);
                    u.w = *reinterpret_cast<uint32_t*>(&t3);
                    reinterpret_cast<uint4*>(dst)[c * 4 + q] = u;
                }
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
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
        if (e != cudaSuccess) {
            snprintf(g_attn_err, sizeof g_attn_err, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured.set();
    }
    if (op.p.seq < ATT_TILE)
        attn_fwd_kernel<64><<<op.grid, ATT_THREADS, ATT_SMEM, st>>>(op.p);
    else
        attn_fwd_kernel<128><<<op.grid, ATT_THREADS, ATT_SMEM, st>>>(op.p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_attn_err, sizeof g_attn_err, "attn launch: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace dxmi
