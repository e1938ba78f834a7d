// Tile epilogue shared by the persistent GEMM kernels (gemm_tc2.cu: one CTA per tile; gemm_tc2p.cu: cta_group::2 pairs).
#pragma once
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace dxmi {

static constexpr int TILE_M = 128;
static constexpr int TILE_K = 64;
static constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
static constexpr int EPI_SLOT_BYTES = 128 * 128;           // 128 rows x 128 bytes

__device__ __forceinline__ float act2(float v, int act) {
    if (act == ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
    if (act == ACT_SILU) return __fdividef(v, 1.f + __expf(-v));
    return v;
}

// ------------------------------------------------------------------------------------------------ tile epilogue
// Per-thread constants of the 8 epilogue warps.
struct EpiCtx {
    uint8_t* slot0;     // 2 staging slots of 128 rows x 128 bytes (32 fp32 columns), SWIZZLE_128B pattern
    float2* sst;        // 2 KB: GroupNorm partials [8 warps][s1 | s2][32 columns] floats, or softmax row stats [128] float2
    uint32_t taddr;     // TMEM address of this warp's lane quarter (column 0 of accumulator 0)
    uint32_t stage_off; // byte offset of this thread's staging row + its 4 swizzled 16-byte units base
    int sw;             // row & 7 (swizzle key) of the staging row this thread writes in phase A
    int hsel;           // which 16 of the chunk's 32 columns this warp stages
    int e, ew;          // epilogue thread / warp index
    int rt0;            // first of the 4 tile rows (rt0 + 4 i) this thread finishes in phase B
    int bu;             // 16-byte unit (4 columns) of the staging row this thread finishes
    bool brs0;          // lane owns the per-warp statistics write
};

enum EpiMode { EPI_BIAS = 0, EPI_ROWVEC = 1, EPI_RESIDUAL = 2, EPI_GENERIC = 3 };

// Drains one 128 x BLOCK_N accumulator tile.
//   Phase A: raw fp32 accumulators TMEM -> swizzled staging slot. Warp w may only read TMEM lanes 32*(w%4)..+31; the
//            two warps that share a lane quarter split the chunk's 32 columns.
//   Phase B: 8 lanes <-> one 128-byte staging row: alpha / bias / row vector / residual / activation / softmax,
//            conversion, fully coalesced global stores, GroupNorm partial sums - on operands prefetched one chunk
//            ahead with coalesced loads.  Staging is double buffered: one named barrier per chunk.
// MODE selects straight-line code for the three hot operator shapes of the U-Nets (bias only, + per-image row vector,
// + residual); EPI_GENERIC keeps every switch at run time (softmax, activations, per-row bias, fp32 output).
// (Measured, profiles/r01_cta_timeline_*: the epilogue is issue / latency bound - 2 warps per scheduler - so code
//  size and dependent chains matter more than bytes.)
template <int V>
struct EpiSlot {
    static constexpr int value = V;
};

// Epilogue operands prefetched ONE TILE AHEAD (a rolling window of four 32-column chunks that crosses tile boundaries).
// Why a whole tile: the SM's 64 B/clk return port is shared with the producer's TMA stages - with 100-150 KB of operand tiles
// queued, ANY load the epilogue issues takes 1.3-2 us (tools/bench_n128.py: a 3-deep instead of a 4-deep operand ring made the
// row-vector / residual epilogues 5-8 % FASTER).  One chunk of lead (round 1) or the start of the tile (early round 2) is not enough.
// Measured (tools/bench_n128.py, 128-ch 3x3 conv, B=256): rolling the row vector one tile ahead 86.6 -> 80.8 us; rolling the residual
// words 85.7 -> 88.6 us (64 carried registers crowd the drain loop), so the residual keeps its one-chunk-ahead prefetch.
#ifndef EPI_ROLL_RESIDUAL
#define EPI_ROLL_RESIDUAL 1
#endif
// RW = window size in 32-column chunks; it must divide the chunks of a tile: 4 for 128- / 256-wide tiles, 3 for 192-wide tiles (6 chunks -
// every ImageNet-64 / LSUN width; without a window their operands fell back to one chunk of lead, and the residual epilogue of the
// short-K 1x1 projections ran at ~10 us per tile: profiles/r02 gemm table, 65536 x 384 x 384 at 0.27 PFLOP/s), 2 for 64-wide tiles.
template <int MODE, int RW = 4>
struct EpiCarry {  // (templated so that each epilogue shape carries only its own operands across tiles)
    uint2 res[(MODE == 2 /*EPI_RESIDUAL*/ && EPI_ROLL_RESIDUAL) ? RW : 1][4];                    // [slot][row]: residual words
    float4 rv[MODE == 1 /*EPI_ROWVEC*/ ? RW : 1];                         // [slot]: per-image row vector of a tile inside one image
    float4 bias[(MODE == 1 || (MODE == 2 && EPI_ROLL_RESIDUAL)) ? RW : 1];                       // [slot]
    int tile_key;                                                        // which tile the slots were primed for (-1: none)
#ifdef DXMI_EPI_PROFILE
    long long prof[8];  // cycles: acc wait | first stage + barrier | stage next | finish + store | prefetch + fold | barrier 2 | publish | barrier 1
    long long tprev;
#endif
};
#ifdef DXMI_EPI_PROFILE
#define EPI_T0() cy.tprev = clock64()
#define EPI_T(k) { const long long now_ = clock64(); cy.prof[k] += now_ - cy.tprev; cy.tprev = now_; }
#else
#define EPI_T0()
#define EPI_T(k)
#endif
__device__ __forceinline__ int epi_tile_key(int m_tile, int col0, int batch) { return (m_tile * 31 + (col0 >> 5)) * 7 + batch; }

// `acc_full` / `acc_parity`: the accumulator-ready barrier of this tile.  The epilogue issues its operand prefetches (the
// residual up to four 32-column chunks ahead: ncu round 2 showed the N = 128 convolutions epilogue bound on exactly these
// L2-latency loads) BEFORE it waits for the accumulator, so they fly during the tile's main loop.
template <int MODE, bool STATS, int RW = 4>
__device__ __forceinline__ void epi_tile(const ConvGemmParams& p, const EpiCtx& cx, uint32_t tacc_col, uint32_t tmem_empty_addr,
                                         int m_tile, int col0, int nch, int batch, uint32_t& out_cnt, uint64_t* acc_full,
                                         uint32_t acc_parity, EpiCarry<MODE, RW>& cy, int next_m_tile = -1, int next_col0 = 0, int next_batch = 0) {
    const int row0 = m_tile * TILE_M;
    constexpr int CH = 32;
    const int n_total = p.N_total;
    const float alpha = p.alpha;
    const bool g_res = MODE == EPI_GENERIC && p.residual != nullptr;
    const bool g_rv = MODE == EPI_GENERIC && p.rowvec != nullptr;
    const bool g_bm = MODE == EPI_GENERIC && p.bias != nullptr && p.bias_along_m;
    const bool g_sm = MODE == EPI_GENERIC && p.softmax != 0;
    const bool g_f32 = MODE == EPI_GENERIC && p.out_fp32 != 0;
    const bool g_gate = MODE == EPI_GENERIC && p.gate != nullptr;
    const int g_act = MODE == EPI_GENERIC ? p.act : ACT_NONE;
    const bool has_bias = p.bias != nullptr && !(MODE == EPI_GENERIC && p.bias_along_m);
    const bool use_rv = MODE == EPI_ROWVEC || g_rv;
    const bool use_res = MODE == EPI_RESIDUAL || g_res;

    // ---- per-tile invariants: the 4 rows this thread finishes
    bool rok[4];
    long long ooff[4], roff[4], goff[4];
    const float* rvp[4];
    float bm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long r;
        long long img;
        bool ok;
        if (p.halo) {
            // halo tile: row i <-> position p0 + i of the zero-padded (W+2)-wide grid of image m_tile / tpi
            const int Wp = p.halo_W + 2;
            img = m_tile / p.halo_tpi;
            const int pos = (m_tile - static_cast<int>(img) * p.halo_tpi) * TILE_M + cx.rt0 + 4 * i;
            const int hh = pos / Wp, ww = pos - hh * Wp;
            ok = ww < p.halo_W && hh < p.halo_H;
            r = (img * p.halo_H + hh) * p.halo_W + ww;
        } else {
            const int r32 = row0 + cx.rt0 + 4 * i;  // rows are ints (M_total is); a 64-bit division here cost ~1 us per tile
            r = r32;
            ok = r32 < p.M_total;
            img = (use_rv && !((p.rows_per_image & 127) == 0)) ? r32 / p.rows_per_image : 0;
        }
        rok[i] = ok && p.dbg_mode != 2;
        // up2 mode: tile row (n, y, x) of phase `batch` = (py, px) -> output pixel (n, 2y + py, 2x + px) = 4 r - 2 x + py 2W + px
        const long long r_out = p.up2 ? 4 * r - 2 * (r & p.up2_wmask) + (batch >> 1) * p.up2_w2 + (batch & 1) : r;
        ooff[i] = batch * p.out_batch_stride + r_out * p.ldo;
        roff[i] = use_res ? batch * p.res_batch_stride + r * p.ldr : 0;
        goff[i] = g_gate ? r * p.ldg : 0;
        rvp[i] = (use_rv && rok[i]) ? p.rowvec + img * p.ldrv : nullptr;
        bm[i] = (g_bm && rok[i]) ? __ldg(p.bias + r) : 0.f;
    }
    // per-tile base pointers at this thread's first column (the ncu source view charged 8 % of the kernel's instructions to
    // re-deriving 64-bit addresses per chunk, and 7 % to an integer division inside the row-vector prefetch)
    const int cc0 = col0 + cx.bu * 4;
    char* optr[4];
    const __nv_bfloat16* rptr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        optr[i] = reinterpret_cast<char*>(p.out) + (ooff[i] + cc0) * (g_f32 ? 4 : 2);
        rptr[i] = use_res ? p.residual + roff[i] + cc0 : nullptr;
    }
    float4 pf_bias = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pf_rv[4];
    uint2 pf_res[4], pf_gate[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pf_rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pf_gate[i] = make_uint2(0u, 0u);
        pf_res[i] = make_uint2(0u, 0u);
    }
    // a 128-row tile of a map with >= 128 pixels lies inside ONE image: the per-image row vector (the ResBlock's time-embedding
    // projection) is then just a second bias - one load and no per-row adds
    const bool rv_uniform = use_rv && !p.halo && (p.rows_per_image & 127) == 0;
    // (a tile past the last row - odd tail of a CTA pair - reads the last image's vector: masked at the store, but never out of bounds)
    const int rv_last = rv_uniform ? (p.M_total - 1) / p.rows_per_image : 0;
    const float* rv_base = rv_uniform ? p.rowvec + static_cast<long long>(min(row0 / p.rows_per_image, rv_last)) * p.ldrv : nullptr;
    // ---- rolling one-tile-ahead window (hot modes, RW | nch): slot k = chunk % RW
    const bool roll = (MODE == EPI_ROWVEC || (MODE == EPI_RESIDUAL && EPI_ROLL_RESIDUAL)) && (nch % RW) == 0 && !p.halo && (MODE != EPI_ROWVEC || rv_uniform);
    const bool have_next = next_m_tile >= 0;
    const __nv_bfloat16* rptr_n[4];
    bool rok_n[4];
    const float* rv_base_n = nullptr;
    int cc0_n = 0;
    if (roll && have_next) {
        const int nrow0 = next_m_tile * TILE_M;
        cc0_n = next_col0 + cx.bu * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r32 = nrow0 + cx.rt0 + 4 * i;
            rok_n[i] = r32 < p.M_total;
            rptr_n[i] = MODE == EPI_RESIDUAL ? p.residual + next_batch * p.res_batch_stride + static_cast<long long>(r32) * p.ldr + cc0_n : nullptr;
        }
        if (MODE == EPI_ROWVEC) rv_base_n = p.rowvec + static_cast<long long>(min(nrow0 / p.rows_per_image, rv_last)) * p.ldrv;
    }
    // load chunk `tc` of THIS tile (next == false) or of the NEXT tile into slot k
    auto roll_load = [&](auto slot_c, int tc, bool next) {
        constexpr int k = (MODE == EPI_ROWVEC || (MODE == EPI_RESIDUAL && EPI_ROLL_RESIDUAL)) ? decltype(slot_c)::value : 0;
        const int pcc = (next ? cc0_n : cc0) + tc * CH;
        if (pcc >= n_total) return;
        if (has_bias) cy.bias[k] = __ldg(reinterpret_cast<const float4*>(p.bias + pcc));
        if (MODE == EPI_ROWVEC) cy.rv[MODE == EPI_ROWVEC ? k : 0] = __ldg(reinterpret_cast<const float4*>((next ? rv_base_n : rv_base) + pcc));
        if (MODE == EPI_RESIDUAL) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool okr = next ? rok_n[i] : rok[i];
                if (okr) cy.res[(MODE == EPI_RESIDUAL && EPI_ROLL_RESIDUAL) ? k : 0][i] = __ldg(reinterpret_cast<const uint2*>((next ? rptr_n[i] : rptr[i]) + tc * CH));
            }
        }
    };
    auto prefetch = [&](int pcc) {  // non-rolling modes: one chunk ahead
        const bool pok = pcc < n_total;
        if (has_bias && pok) pf_bias = __ldg(reinterpret_cast<const float4*>(p.bias + pcc));
        if (use_rv) {
            if (rv_uniform) {
                if (pok) pf_rv[0] = __ldg(reinterpret_cast<const float4*>(rv_base + pcc));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (rvp[i] && pok) pf_rv[i] = __ldg(reinterpret_cast<const float4*>(rvp[i] + pcc));
            }
        }
        if (use_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (rok[i] && pok) pf_res[i] = __ldg(reinterpret_cast<const uint2*>(rptr[i] + (pcc - cc0)));
        }
        if (g_gate) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (rok[i] && pok) pf_gate[i] = __ldg(reinterpret_cast<const uint2*>(p.gate + goff[i] + pcc));
        }
    };
    if (roll) {
        if (cy.tile_key != epi_tile_key(m_tile, col0, batch)) {  // first tile of this CTA: prime the window
            roll_load(EpiSlot<0>{}, 0, false);
            roll_load(EpiSlot<1>{}, 1, false);
            if constexpr (RW > 2) roll_load(EpiSlot<2>{}, 2, false);
            if constexpr (RW > 3) roll_load(EpiSlot<3>{}, 3, false);
        }
        cy.tile_key = have_next ? epi_tile_key(next_m_tile, next_col0, next_batch) : -1;
    } else {
        prefetch(cc0);
    }

    // ---- accumulator ready?
    EPI_T0();
    ptx::mbar_wait(acc_full, acc_parity);
    ptx::tc_fence_after();
    EPI_T(0);
    if (g_sm) {
        // row max / sum over the whole accumulator row (N_total == BLOCK_N): one warp per lane quarter
        if (cx.hsel == 0) {
            float sm_max = -INFINITY, sm_sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < n_total; c += 32) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(cx.taddr + tacc_col + c, v);
                ptx::tmem_ld_wait();
                float cmax = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) cmax = fmaxf(cmax, __uint_as_float(v[j]) * alpha);
                const float nmax = fmaxf(sm_max, cmax);
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) part += __expf(__uint_as_float(v[j]) * alpha - nmax);
                sm_sum = sm_sum * __expf(sm_max - nmax) + part;
                sm_max = nmax;
            }
            // staging row index of this thread == its accumulator row
            cx.sst[(cx.stage_off >> 10) * 8 + ((cx.stage_off >> 7) & 7)] = make_float2(sm_max, 1.f / sm_sum);
        }
        ptx::named_bar_sync(2, 256);
    }

    const bool stat_direct = STATS && !p.halo && p.stats_seg == 16;

    // ---- phase A: 16 accumulator columns of this warp's 32 rows -> staging slot of chunk c
    const uint32_t cnt0 = out_cnt;
    // tcgen05.ld is asynchronous: the loads of chunk c + 1 are ISSUED before chunk c is finished and waited for after it, so the
    // TMEM read latency (~500 cycles next to a running MMA: tools/epi_profile.py) hides behind phase B instead of adding to it
    auto stage_issue = [&](int c, uint32_t (&v)[16]) { ptx::tmem_ld_32x32b_x16(cx.taddr + tacc_col + (c * CH + cx.hsel * 16), v); };
    auto stage_finish = [&](int c, uint32_t (&v)[16]) {
        uint8_t* slot = cx.slot0 + ((cnt0 + c) & 1) * EPI_SLOT_BYTES;
        ptx::tmem_ld_wait();
        if (c == nch - 1) {
            // every accumulator column this warp stages is now in registers: hand the TMEM buffer back
            ptx::tc_fence_before();
            // (shared::cluster address: the CTA's own barrier, or the leader CTA's in a cta_group::2 pair)
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tmem_empty_addr) : "memory");
        }
        uint8_t* srow = slot + cx.stage_off;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(srow + (((cx.hsel * 4 + q) ^ cx.sw) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    };
    // One barrier per chunk: chunk c+1 is staged (other slot) BEFORE the barrier that ends chunk c, which orders both
    // "c+1 staged" and "slot of c free".
    {
        uint32_t v0[16];
        stage_issue(0, v0);
        stage_finish(0, v0);
    }
    ptx::named_bar_sync(1, 256);
    EPI_T(1);
    // masks only where a tile can be ragged (uniform branch; the statistics add 18 instructions per row otherwise)
    const bool tile_full = row0 + TILE_M <= p.M_total && col0 + nch * CH <= n_total && p.dbg_mode != 2 && !p.halo;
    const int lg = (threadIdx.x & 31) >> 3;  // lane's row sub-index

    auto do_chunk = [&](int c, auto slot_c) {
        constexpr int rs = (MODE == EPI_ROWVEC || (MODE == EPI_RESIDUAL && EPI_ROLL_RESIDUAL)) ? decltype(slot_c)::value : 0;  // rolling-window slot (chunk & 3)
        const int col = col0 + c * CH;
        const int cc = col + cx.bu * 4;
        const bool col_ok = cc < n_total;
        uint8_t* slot = cx.slot0 + (out_cnt & 1) * EPI_SLOT_BYTES;
        // (the generic shape carries too many operands to hold 16 more registers across phase B: it issues the load late)
        constexpr bool AHEAD = MODE != EPI_GENERIC;
        uint32_t vn[16];
        if (AHEAD && c + 1 < nch) stage_issue(c + 1, vn);
        EPI_T(2);

        // ---- phase B
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rt = cx.rt0 + 4 * i;
            const uint4 u = *reinterpret_cast<const uint4*>(slot + (rt >> 3) * 1024 + (rt & 7) * 128 + ((cx.bu ^ (rt & 7)) << 4));
            float x[4];
            if (g_sm) {
                const float2 ms = cx.sst[rt];
                x[0] = __expf(__uint_as_float(u.x) * alpha - ms.x) * ms.y;
                x[1] = __expf(__uint_as_float(u.y) * alpha - ms.x) * ms.y;
                x[2] = __expf(__uint_as_float(u.z) * alpha - ms.x) * ms.y;
                x[3] = __expf(__uint_as_float(u.w) * alpha - ms.x) * ms.y;
            } else {
                float4 ad = roll ? cy.bias[rs] : pf_bias;
                if (!has_bias) ad = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rv_uniform) {
                    const float4 rvv = roll ? cy.rv[MODE == EPI_ROWVEC ? rs : 0] : pf_rv[0];
                    ad.x += rvv.x;
                    ad.y += rvv.y;
                    ad.z += rvv.z;
                    ad.w += rvv.w;
                }
                if (g_bm) {
                    ad.x += bm[i];
                    ad.y += bm[i];
                    ad.z += bm[i];
                    ad.w += bm[i];
                }
                if (use_rv && !rv_uniform) {
                    ad.x += pf_rv[i].x;
                    ad.y += pf_rv[i].y;
                    ad.z += pf_rv[i].z;
                    ad.w += pf_rv[i].w;
                }
                x[0] = fmaf(__uint_as_float(u.x), alpha, ad.x);
                x[1] = fmaf(__uint_as_float(u.y), alpha, ad.y);
                x[2] = fmaf(__uint_as_float(u.z), alpha, ad.z);
                x[3] = fmaf(__uint_as_float(u.w), alpha, ad.w);
                if (use_res) {
                    const uint2 rw = (MODE == EPI_RESIDUAL && roll) ? cy.res[(MODE == EPI_RESIDUAL && EPI_ROLL_RESIDUAL) ? rs : 0][i] : pf_res[i];
                    x[0] += __uint_as_float(rw.x << 16);
                    x[1] += __uint_as_float(rw.x & 0xffff0000u);
                    x[2] += __uint_as_float(rw.y << 16);
                    x[3] += __uint_as_float(rw.y & 0xffff0000u);
                }
                if (g_act != ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = act2(x[j], g_act);
                }
                if (g_gate) {
                    // sign of the saved bf16 activation: positive -> slope 1, else 0.2
                    x[0] *= __uint_as_float(pf_gate[i].x << 16) > 0.f ? 1.f : 0.2f;
                    x[1] *= __uint_as_float(pf_gate[i].x & 0xffff0000u) > 0.f ? 1.f : 0.2f;
                    x[2] *= __uint_as_float(pf_gate[i].y << 16) > 0.f ? 1.f : 0.2f;
                    x[3] *= __uint_as_float(pf_gate[i].y & 0xffff0000u) > 0.f ? 1.f : 0.2f;
                }
            }
            const bool ok = rok[i] && col_ok;
            if (g_f32) {
                if (ok) *reinterpret_cast<float4*>(optr[i] + static_cast<long long>(c) * (CH * 4)) = make_float4(x[0], x[1], x[2], x[3]);
            } else {
                __nv_bfloat162 b0 = __floats2bfloat162_rn(x[0], x[1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(x[2], x[3]);
                uint32_t w0 = *reinterpret_cast<uint32_t*>(&b0), w1 = *reinterpret_cast<uint32_t*>(&b1);
                if (ok) *reinterpret_cast<uint2*>(optr[i] + c * (CH * 2)) = make_uint2(w0, w1);
                if (STATS) {
                    // statistics of the values the consumer will read (bf16-rounded); masked rows / columns add zero
                    if (!tile_full) {
                        w0 = ok ? w0 : 0u;
                        w1 = ok ? w1 : 0u;
                    }
                    const float a0 = __uint_as_float(w0 << 16), a1 = __uint_as_float(w0 & 0xffff0000u);
                    const float a2 = __uint_as_float(w1 << 16), a3 = __uint_as_float(w1 & 0xffff0000u);
                    s1[0] += a0; s2[0] = fmaf(a0, a0, s2[0]);
                    s1[1] += a1; s2[1] = fmaf(a1, a1, s2[1]);
                    s1[2] += a2; s2[2] = fmaf(a2, a2, s2[2]);
                    s1[3] += a3; s2[3] = fmaf(a3, a3, s2[3]);
                }
            }
        }
        EPI_T(3);
        if (roll) {
            // this slot is consumed: refill it with the chunk RW ahead - same tile, or the NEXT tile's chunk (c + RW - nch)
            if (c + RW < nch) roll_load(slot_c, c + RW, false);
            else if (have_next) roll_load(slot_c, c + RW - nch, true);
        } else if (c + 1 < nch) {
            prefetch(cc + CH);  // next chunk's operands fly during the stats tail and phase A
        }
        // this warp's 16 rows: fold the 4 row sub-indices (lanes xor 8, 16).  Halving butterfly: each round a lane hands half of
        // its values to its partner and keeps the partner's other half - 6 shuffles instead of 16 (the SM shuffles one warp per
        // clock) with the SAME association (a0 + a1) + (a2 + a3).  Afterwards lane sub-index g owns, for its column quad:
        // g = 0: s1 of columns 0,1 | g = 1: s2 of columns 0,1 | g = 2: s1 of columns 2,3 | g = 3: s2 of columns 2,3
        float u0 = 0.f, u1 = 0.f;
        if (STATS) {
            const bool odd = lg & 1, hi = lg >> 1;
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float keep = odd ? s2[j] : s1[j], send = odd ? s1[j] : s2[j];
                t[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            const float k0 = hi ? t[2] : t[0], k1 = hi ? t[3] : t[1];
            const float d0 = hi ? t[0] : t[2], d1 = hi ? t[1] : t[3];
            u0 = k0 + __shfl_xor_sync(0xffffffffu, d0, 16);
            u1 = k1 + __shfl_xor_sync(0xffffffffu, d1, 16);
        }
        if (c + 1 < nch) {
            if (!AHEAD) stage_issue(c + 1, vn);
            stage_finish(c + 1, vn);
        }
        EPI_T(4);
        float* sstf = reinterpret_cast<float*>(cx.sst);  // [8 warps][2: s1 | s2][32 columns]
        if (STATS && stat_direct) {
            // 16-row segments: a segment is exactly the 16 rows of this warp - publish straight to global memory (no shared
            // memory round trip, no second barrier; the consumer's finalize sums HW/16 partials per image in a fixed order)
            const int srow = row0 + ((cx.ew & 3) * 2 + (cx.ew >> 2)) * 16;
            if (col_ok && srow < p.M_total) {
                float* dst = p.stats + (static_cast<long long>(srow >> 4) * n_total + cc + (lg >> 1) * 2) * 2 + (lg & 1);
                dst[0] = u0;
                dst[2] = u1;
            }
        } else if (STATS) {
            // hand the warp partials to the publisher warp (epi_publish_tile, an otherwise idle warp of warp group 0): barrier 3
            // = "the previous chunk's partials have been read" (never waits in practice: the publisher has a whole chunk of time),
            // barrier 2 = "this chunk's partials are in shared memory" (arrive only - the epilogue warps do not wait for it)
            ptx::named_bar_sync(3, 288);
            *reinterpret_cast<float2*>(sstf + cx.ew * 64 + (lg & 1) * 32 + cx.bu * 4 + (lg >> 1) * 2) = make_float2(u0, u1);
            ptx::named_bar_arrive(2, 288);
            EPI_T(5);
        }
        EPI_T(6);
        ++out_cnt;
        ptx::named_bar_sync(1, 256);
        EPI_T(7);
    };
    if (roll) {
#pragma unroll 1
        for (int c0 = 0; c0 < nch; c0 += RW) {
            do_chunk(c0, EpiSlot<0>{});
            do_chunk(c0 + 1, EpiSlot<1>{});
            if constexpr (RW > 2) do_chunk(c0 + 2, EpiSlot<2>{});
            if constexpr (RW > 3) do_chunk(c0 + 3, EpiSlot<3>{});
        }
    } else {
#pragma unroll 1
        for (int c = 0; c < nch; ++c) do_chunk(c, EpiSlot<0>{});
    }
}


// Lean drain for bias-only tiles WITHOUT statistics (the q|k|v projections of the ADM attention blocks: M = 65536, N = 1152, K = 384 runs
// 73 us with the staged epilogue against a 31 us HBM / 41 us tensor floor - short-K GEMMs are bound by the epilogue's output rate, about
// 1000 cycles per 128 x 32 chunk, tools/bench_1x1.py).  Nothing has to cross rows here, so there is no reason to transpose through shared
// memory: thread <-> accumulator row, 16 fp32 columns per thread and chunk straight from TMEM (the load of chunk c + 1 is in flight while
// chunk c is converted), alpha / bias, bf16, one 32-byte row segment = two 16-byte stores.  No staging, no named barriers, ~35 instead of
// ~110 instructions per thread and chunk.  The row-per-thread stores touch 32 lines per instruction - what made the round-1 epilogue slow
// with four warps and 16-byte segments - but every sector is written whole by the same thread and eight warps keep the LSU queue full.
// MEASURED NEGATIVE on B200 (tools/bench_1x1.py, option "lean_epi", off by default): M = 65536, K = 384 bias-only, N = 384: 38.5 us vs 29.9 us
// staged; N = 1152: 96.0 vs 72.9 us; ImageNet-64 T = 10: 394.9 vs 400.6 img/s.  A third of the instructions, and still slower: a store
// instruction that touches 32 lines occupies the LSU for 32 cycles however few instructions surround it - the transposing epilogue's
// full-line stores are what the memory pipe wants.
__device__ __forceinline__ bool epi_lean_ok(const ConvGemmParams& p) {
    return p.lean_epi && !p.halo && p.dbg_mode == 0 && (p.N_total & 15) == 0 && (p.ldo & 7) == 0;
}
__device__ __forceinline__ void epi_tile_lean(const ConvGemmParams& p, const EpiCtx& cx, uint32_t tacc_col, uint32_t tmem_empty_addr, int m_tile,
                                              int col0, int nch, int batch, uint64_t* acc_full, uint32_t acc_parity) {
    const int row_in_tile = (cx.stage_off >> 10) * 8 + ((cx.stage_off >> 7) & 7);
    const long long r = static_cast<long long>(m_tile) * TILE_M + row_in_tile;
    const bool ok = r < p.M_total;
    const long long r_out = p.up2 ? 4 * r - 2 * (r & p.up2_wmask) + (batch >> 1) * p.up2_w2 + (batch & 1) : r;
    const int cc0 = col0 + cx.hsel * 16;
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + batch * p.out_batch_stride + r_out * p.ldo + cc0;
    const float alpha = p.alpha;
    const bool has_bias = p.bias != nullptr;
    const int n_total = p.N_total;
    ptx::mbar_wait(acc_full, acc_parity);
    ptx::tc_fence_after();
    uint32_t va[16], vb[16];
    ptx::tmem_ld_32x32b_x16(cx.taddr + tacc_col + cx.hsel * 16, va);
    auto finish = [&](int c, uint32_t (&v)[16]) {
        const int cc = cc0 + c * 32;
        if (cc >= n_total) return;
        float4 b4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) b4[q] = has_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cc) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t w[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x0 = fmaf(__uint_as_float(v[4 * q]), alpha, b4[q].x), x1 = fmaf(__uint_as_float(v[4 * q + 1]), alpha, b4[q].y);
            const float x2 = fmaf(__uint_as_float(v[4 * q + 2]), alpha, b4[q].z), x3 = fmaf(__uint_as_float(v[4 * q + 3]), alpha, b4[q].w);
            __nv_bfloat162 t0 = __floats2bfloat162_rn(x0, x1), t1 = __floats2bfloat162_rn(x2, x3);
            w[2 * q] = *reinterpret_cast<uint32_t*>(&t0);
            w[2 * q + 1] = *reinterpret_cast<uint32_t*>(&t1);
        }
        if (ok) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {
        ptx::tmem_ld_wait();  // chunk c (va)
        if (c + 1 < nch) ptx::tmem_ld_32x32b_x16(cx.taddr + tacc_col + ((c + 1) * 32 + cx.hsel * 16), vb);
        else {
            ptx::tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tmem_empty_addr) : "memory");
        }
        finish(c, va);
        if (c + 1 < nch) {
            ptx::tmem_ld_wait();  // chunk c + 1 (vb)
            if (c + 2 < nch) ptx::tmem_ld_32x32b_x16(cx.taddr + tacc_col + ((c + 2) * 32 + cx.hsel * 16), va);
            else {
                ptx::tc_fence_before();
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tmem_empty_addr) : "memory");
            }
            finish(c + 1, vb);
        }
    }
}

// Statistics publisher (ONE warp, not an epilogue warp): for every chunk of a tile, combines the 8 warp partials per column
// in a fixed order - per row segment: 32-row blocks ascending, the two 16-row halves of each - and writes one (sum, sum of
// squares) pair per segment and column.  Keeps 16 shared-memory loads, the adds and the store off the epilogue warps' critical
// path (tools/epi_profile.py: they cost 450-600 cycles per chunk there, a quarter of the tile's drain time).
// Must be called for exactly the tiles, in the order, that the epilogue warps process when epi_stats_published(p).
__device__ __forceinline__ bool epi_stats_published(const ConvGemmParams& p) {
    return p.stats != nullptr && !p.out_fp32 && (p.halo || p.stats_seg != 16);
}
__device__ __forceinline__ void epi_publish_tile(const ConvGemmParams& p, const float* sstf, int m_tile, int col0, int nch, int lane, int batch = 0) {
    const int row0 = m_tile * TILE_M;
    const int n_total = p.N_total;
    const int stat_seg = !p.halo ? p.stats_seg : 128;
    const int bps = stat_seg >> 5;  // 32-row blocks per segment: 1, 2 or 4
    const int stat_shift = stat_seg == 128 ? 7 : (stat_seg == 64 ? 6 : 5);
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int col = col0 + c * 32 + lane;
        ptx::named_bar_sync(2, 288);
        float2 pb[4][2];
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) pb[b2][hh] = make_float2(sstf[(hh * 4 + b2) * 64 + lane], sstf[(hh * 4 + b2) * 64 + 32 + lane]);
        float2 seg[4];
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) {
            a.x += pb[b2][0].x;
            a.y += pb[b2][0].y;
            a.x += pb[b2][1].x;
            a.y += pb[b2][1].y;
            seg[b2] = a;
            if (((b2 + 1) & (bps - 1)) == 0) a = make_float2(0.f, 0.f);
        }
        // the sums above consumed every loaded value: the partials may be overwritten
        ptx::named_bar_arrive(3, 288);
        if (col < n_total) {
            if (p.halo) {
                *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(m_tile) * n_total + col) * 2) = seg[3];
            } else {
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) {
                    const int srow = row0 + (b2 + 1 - bps) * 32;  // first row of the segment that ends with block b2
                    if (((b2 + 1) & (bps - 1)) == 0 && srow < p.M_total) {
                        int sidx = srow >> stat_shift;
                        if (p.up2) {  // partials of one OUTPUT image stay contiguous: [image][phase][segment of the low-resolution image]
                            const int img = sidx / p.up2_spi;
                            sidx = (img * 4 + batch) * p.up2_spi + (sidx - img * p.up2_spi);
                        }
                        *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(sidx) * n_total + col) * 2) = seg[b2];
                    }
                }
            }
        }
    }
}

}  // namespace dxmi
