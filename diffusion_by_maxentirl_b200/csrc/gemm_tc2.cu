// Persistent tcgen05 implicit-GEMM convolution kernel (v2). Same operator contract as gemm_tc.cu / gemm_tc.cuh.
//
// What changed against v1 (measured with per-CTA globaltimer stamps, profiles/r01_cta_timeline.txt: the row-per-thread
// direct-store epilogue cost 7.5 us per 128x256 tile against a 12 us main loop and nothing overlapped it):
//   * one CTA per SM, looping over output tiles (static round-robin), so barrier init / TMEM alloc / descriptor
//     prefetch are paid once per SM instead of once per tile;
//   * two TMEM accumulators: the MMA warp starts tile i+1 while the epilogue warps drain tile i;
//   * the epilogue is two-phase: accumulator rows go through a swizzled shared-memory staging slot, then 8 lanes per
//     row add the residual (prefetched with coalesced loads), apply the activation and store full 128-byte lines.
//     (A TMA-store epilogue was measured first: its stores queue behind the producer's loads in the TMA unit and cost
//     1.4 us per 16 KB chunk, profiles/r01_cta_timeline_v2.txt);
//   * optional fused GroupNorm partial statistics: per 32-row segment and output column, (sum, sumsq) of the
//     bf16-rounded outputs via a warp reduce-scatter (31 shuffles per 32 columns), so the consumer GroupNorm needs no
//     statistics pass over HBM.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2..9 epilogue
// (warp w may read TMEM lanes 32*(w%4)..+31).
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <cstdio>

namespace dxmi {

static constexpr int TILE_M = 128;
static constexpr int TILE_K = 64;
static constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
static constexpr int NUM_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps
static constexpr int EPI_SLOT_BYTES = 128 * 128;  // 128 rows x 128 bytes

template <int BLOCK_N>
struct Cfg2 {
    static constexpr int B_STAGE_BYTES = BLOCK_N * TILE_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = BLOCK_N > 128 ? 4 : (BLOCK_N > 64 ? 6 : 8);
    static constexpr int ACC_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SM_OUT = RING_BYTES;                       // 2 output staging slots
    static constexpr int SM_STAT = SM_OUT + 2 * EPI_SLOT_BYTES;     // [8 warps][32 columns] float2 GroupNorm partials + softmax row stats
    static constexpr int SM_BAR = SM_STAT + 2048;                   // mbarriers + TMEM slot
    static_assert(SM_BAR + 512 <= 227 * 1024, "shared memory budget");
    static constexpr int SMEM_BYTES = SM_BAR + 512;
};

__device__ __forceinline__ long long gtimer2() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DBG2(slot) \
    if (p.dbg_times) p.dbg_times[(long long)blockIdx.x * 8 + (slot)] = gtimer2();
// cycle-resolution stamps of the epilogue leader for chunk `cidx` of the CTA's second tile (steady state)


__device__ __forceinline__ float act2(float v, int act) {
    if (act == ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
    if (act == ACT_SILU) return __fdividef(v, 1.f + __expf(-v));
    return v;
}

// ------------------------------------------------------------------------------------------------ tile epilogue
// Per-thread constants of the 8 epilogue warps.
struct EpiCtx {
    uint8_t* slot0;     // 2 staging slots of 128 rows x 128 bytes (32 fp32 columns), SWIZZLE_128B pattern
    float2* sst;        // [8 warps][32 columns] GroupNorm partials / softmax row stats
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
template <int MODE, bool STATS>
__device__ __forceinline__ void epi_tile(const ConvGemmParams& p, const EpiCtx& cx, uint32_t tacc_col, uint64_t* tmem_empty_bar,
                                         int m_tile, int col0, int nch, int batch, uint32_t& out_cnt) {
    const int row0 = m_tile * TILE_M;
    constexpr int CH = 32;
    const int n_total = p.N_total;
    const float alpha = p.alpha;
    const bool g_res = MODE == EPI_GENERIC && p.residual != nullptr;
    const bool g_rv = MODE == EPI_GENERIC && p.rowvec != nullptr;
    const bool g_bm = MODE == EPI_GENERIC && p.bias != nullptr && p.bias_along_m;
    const bool g_sm = MODE == EPI_GENERIC && p.softmax != 0;
    const bool g_f32 = MODE == EPI_GENERIC && p.out_fp32 != 0;
    const int g_act = MODE == EPI_GENERIC ? p.act : ACT_NONE;
    const bool has_bias = p.bias != nullptr && !(MODE == EPI_GENERIC && p.bias_along_m);
    const bool use_rv = MODE == EPI_ROWVEC || g_rv;
    const bool use_res = MODE == EPI_RESIDUAL || g_res;

    // ---- per-tile invariants: the 4 rows this thread finishes
    bool rok[4];
    long long ooff[4], roff[4];
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
            r = static_cast<long long>(row0) + cx.rt0 + 4 * i;
            ok = r < p.M_total;
            img = use_rv ? r / p.rows_per_image : 0;
        }
        rok[i] = ok && p.dbg_mode != 2;
        ooff[i] = batch * p.out_batch_stride + r * p.ldo;
        roff[i] = use_res ? batch * p.res_batch_stride + r * p.ldr : 0;
        rvp[i] = (use_rv && rok[i]) ? p.rowvec + img * p.ldrv : nullptr;
        bm[i] = (g_bm && rok[i]) ? __ldg(p.bias + r) : 0.f;
    }
    float4 pf_bias = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pf_rv[4];
    uint2 pf_res[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pf_rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pf_res[i] = make_uint2(0u, 0u);
    }
    auto prefetch = [&](int pcc) {
        const bool pok = pcc < n_total;
        if (has_bias && pok) pf_bias = __ldg(reinterpret_cast<const float4*>(p.bias + pcc));
        if (use_rv) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (rvp[i] && pok) pf_rv[i] = __ldg(reinterpret_cast<const float4*>(rvp[i] + pcc));
        }
        if (use_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (rok[i] && pok) pf_res[i] = __ldg(reinterpret_cast<const uint2*>(p.residual + roff[i] + pcc));
        }
    };
    prefetch(col0 + cx.bu * 4);

    // ---- accumulator ready?
    // (the caller has already waited on tmem_full and fenced)
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

    const int stat_seg = (STATS && !p.halo) ? p.stats_seg : 128;  // halo tiles: one partial per tile (tiles never span images)
    const int stat_nseg = 128 / stat_seg, stat_bps = stat_seg >> 5;
    const int stat_shift = stat_seg == 128 ? 7 : (stat_seg == 64 ? 6 : 5);

#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int col = col0 + c * CH;
        const int cc = col + cx.bu * 4;
        const bool col_ok = cc < n_total;
        uint8_t* slot = cx.slot0 + (out_cnt & 1) * EPI_SLOT_BYTES;

        // ---- phase A: 16 accumulator columns of this warp's 32 rows -> staging
        {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(cx.taddr + tacc_col + (c * CH + cx.hsel * 16), v);
            ptx::tmem_ld_wait();
            if (c == nch - 1) {
                // every accumulator column this warp stages is now in registers: hand the TMEM buffer back
                ptx::tc_fence_before();
                ptx::mbar_arrive(tmem_empty_bar);
            }
            uint8_t* srow = slot + cx.stage_off;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(srow + (((cx.hsel * 4 + q) ^ cx.sw) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        ptx::named_bar_sync(1, 256);

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
                float4 ad = pf_bias;
                if (g_bm) {
                    ad.x += bm[i];
                    ad.y += bm[i];
                    ad.z += bm[i];
                    ad.w += bm[i];
                }
                if (use_rv) {
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
                    x[0] += __uint_as_float(pf_res[i].x << 16);
                    x[1] += __uint_as_float(pf_res[i].x & 0xffff0000u);
                    x[2] += __uint_as_float(pf_res[i].y << 16);
                    x[3] += __uint_as_float(pf_res[i].y & 0xffff0000u);
                }
                if (g_act != ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = act2(x[j], g_act);
                }
            }
            const bool ok = rok[i] && col_ok;
            if (g_f32) {
                if (ok) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + ooff[i] + cc) = make_float4(x[0], x[1], x[2], x[3]);
            } else {
                __nv_bfloat162 b0 = __floats2bfloat162_rn(x[0], x[1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(x[2], x[3]);
                uint32_t w0 = *reinterpret_cast<uint32_t*>(&b0), w1 = *reinterpret_cast<uint32_t*>(&b1);
                if (ok) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff[i] + cc) = make_uint2(w0, w1);
                if (STATS) {
                    // statistics of the values the consumer will read (bf16-rounded); masked rows / columns add zero
                    w0 = ok ? w0 : 0u;
                    w1 = ok ? w1 : 0u;
                    const float a0 = __uint_as_float(w0 << 16), a1 = __uint_as_float(w0 & 0xffff0000u);
                    const float a2 = __uint_as_float(w1 << 16), a3 = __uint_as_float(w1 & 0xffff0000u);
                    s1[0] += a0; s2[0] = fmaf(a0, a0, s2[0]);
                    s1[1] += a1; s2[1] = fmaf(a1, a1, s2[1]);
                    s1[2] += a2; s2[2] = fmaf(a2, a2, s2[2]);
                    s1[3] += a3; s2[3] = fmaf(a3, a3, s2[3]);
                }
            }
        }
        if (c + 1 < nch) prefetch(cc + CH);  // next chunk's operands fly during the stats tail and phase A
        if (STATS) {
            // this warp's 16 rows: fold the 4 row sub-indices (lanes xor 8, 16); then the warps that share a row segment
            // are combined through smem in a fixed order and one partial per segment is published
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
            }
            if (cx.brs0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) cx.sst[cx.ew * 32 + cx.bu * 4 + j] = make_float2(s1[j], s2[j]);
            }
            ptx::named_bar_sync(2, 256);
            // (sst is single buffered: the next chunk writes it after its named barrier 1, which every publisher of this
            //  chunk reaches only after it has read sst)
            const int sg = cx.e >> 5, j = cx.e & 31;  // thread publishes column j of segment sg
            if (sg < stat_nseg) {
                float2 a = make_float2(0.f, 0.f);
                for (int b2 = sg * stat_bps; b2 < (sg + 1) * stat_bps; ++b2) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float2 b = cx.sst[(hh * 4 + b2) * 32 + j];
                        a.x += b.x;
                        a.y += b.y;
                    }
                }
                const int srow = row0 + sg * stat_seg;
                if (p.halo) {
                    if (col + j < n_total) *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(m_tile) * n_total + col + j) * 2) = a;
                } else if (col + j < n_total && srow < p.M_total) {
                    *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(srow >> stat_shift) * n_total + col + j) * 2) = a;
                }
            }
        }
        ++out_cnt;
    }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_gemm2_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = Cfg2<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SM_BAR);
    constexpr int MAXB = 12;                      // ring depth upper bound (halo mode re-carves the ring at run time)
    uint64_t* full_bar = bars;                    // [MAXB]  (A+B stages; B stages in halo mode)
    uint64_t* empty_bar = bars + MAXB;            // [MAXB]
    uint64_t* tmem_full = bars + 2 * MAXB;        // [2]
    uint64_t* tmem_empty = bars + 2 * MAXB + 2;   // [2]
    uint64_t* afull_bar = bars + 2 * MAXB + 4;    // [2]  halo A stages
    uint64_t* aempty_bar = bars + 2 * MAXB + 6;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAXB + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles * p.batch_count;
    if (threadIdx.x == 0) { DBG2(0); }

    int k_iters = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s)
        if (s < p.nseg) k_iters += p.seg[s].ntaps * p.seg[s].nchunks;

    if (threadIdx.x == 0) {
        if (ptx::smem_u32(smem) & 1023u) {
            printf("dxmi conv_gemm2: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tmap(&p.a_map[0]);
        if (p.nseg > 1) ptx::prefetch_tmap(&p.a_map[1]);
        if (p.nseg > 2) ptx::prefetch_tmap(&p.a_map[2]);
        ptx::prefetch_tmap(&p.b_map);
        for (int s = 0; s < MAXB; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 256);
            ptx::mbar_init(&afull_bar[s], 1);
            ptx::mbar_init(&aempty_bar[s], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) { DBG2(1); }

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (p.halo) {
            if (ptx::elect_one()) {
                const int Wp = p.halo_W + 2;
                const uint32_t a_bytes = static_cast<uint32_t>(p.halo_rows) * Wp * 128u;
                const int SB = p.halo_sb;
                uint8_t* sB0 = smem + 2 * p.halo_a_stage;
                uint32_t ia = 0, ib = 0;
                for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                    const int n_tile = t % p.n_tiles;
                    const int m_tile = t / p.n_tiles;
                    const int img = m_tile / p.halo_tpi;
                    const int p0 = (m_tile - img * p.halo_tpi) * TILE_M;
                    const int h_first = p0 / Wp;
                    int kbase = 0;
                    for (int s = 0; s < p.nseg; ++s) {
                        const GemmSeg sg = p.seg[s];
                        const CUtensorMap* amap = &p.a_map[sg.map];
                        const int cseg = sg.nchunks * TILE_K;
                        for (int ch = 0; ch < sg.nchunks; ++ch, ++ia) {
                            const uint32_t sa = ia & 1;
                            ptx::mbar_wait(&aempty_bar[sa], ((ia >> 1) & 1) ^ 1);
                            ptx::mbar_expect_tx(&afull_bar[sa], a_bytes);
                            ptx::tma_load_4d(smem + sa * p.halo_a_stage, amap, &afull_bar[sa], ch * TILE_K, -1, h_first - 1, img);
                            for (int tap = 0; tap < sg.ntaps; ++tap, ++ib) {
                                const uint32_t sb = ib % SB;
                                ptx::mbar_wait(&empty_bar[sb], ((ib / SB) & 1) ^ 1);
                                ptx::mbar_expect_tx(&full_bar[sb], Cfg::B_STAGE_BYTES);
                                ptx::tma_load_3d(sB0 + sb * Cfg::B_STAGE_BYTES, &p.b_map, &full_bar[sb], kbase + tap * cseg + ch * TILE_K,
                                                 n_tile * BLOCK_N, 0);
                            }
                        }
                        kbase += sg.ntaps * cseg;
                    }
                }
            }
        } else if (ptx::elect_one()) {
            const int tiles_per_nblk = p.tiles_w * p.tiles_h;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int n_tile = t % p.n_tiles;
                const int mt = t / p.n_tiles;
                const int m_tile = mt % p.m_tiles;
                const int batch = mt / p.m_tiles;
                const int n_blk = m_tile / tiles_per_nblk;
                const int rem = m_tile - n_blk * tiles_per_nblk;
                const int h_blk = rem / p.tiles_w;
                const int w_blk = rem - h_blk * p.tiles_w;
                const int w0 = w_blk * p.bw * p.stride;
                const int h0 = h_blk * p.bh * p.stride;
                const int n0 = p.a_batched ? batch : n_blk * p.bn;
                const int bcoord_n = n_tile * BLOCK_N;
                const int bcoord_b = p.b_batched ? batch : 0;
                int kk = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const GemmSeg sg = p.seg[s];
                    const CUtensorMap* amap = &p.a_map[sg.map];
                    for (int tap = 0; tap < sg.ntaps; ++tap) {
                        const int r = (sg.ntaps == 9) ? tap / 3 : 0;
                        const int q = (sg.ntaps == 9) ? tap - 3 * r : 0;
                        for (int ch = 0; ch < sg.nchunks; ++ch, ++it, ++kk) {
                            const uint32_t stage = it % STAGES;
                            const uint32_t ph = (it / STAGES) & 1;
                            ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
                            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                            uint8_t* sb = sa + A_STAGE_BYTES;
                            ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                            ptx::tma_load_4d(sa, amap, &full_bar[stage], ch * TILE_K, w0 + q - sg.pad, h0 + r - sg.pad, n0);
                            ptx::tma_load_3d(sb, &p.b_map, &full_bar[stage], kk * TILE_K, bcoord_n, bcoord_b);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (p.halo) {
            if (ptx::elect_one()) {
                constexpr uint32_t idesc = ptx::make_idesc(/*bf16*/ 1, TILE_M, BLOCK_N);
                const int Wp = p.halo_W + 2;
                const int SB = p.halo_sb;
                const uint32_t sB0 = ptx::smem_u32(smem + 2 * p.halo_a_stage);
                uint32_t ia = 0, ib = 0, ti = 0;
                for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
                    const int m_tile = t / p.n_tiles;
                    const int p0 = (m_tile % p.halo_tpi) * TILE_M;
                    const int w0 = p0 % Wp;
                    const uint32_t acc = ti & 1;
                    ptx::mbar_wait(&tmem_empty[acc], ((ti >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tacc = tmem_base + acc * Cfg::ACC_COLS;
                    uint32_t first = 1;
                    for (int s = 0; s < p.nseg; ++s) {
                        const GemmSeg sg = p.seg[s];
                        for (int ch = 0; ch < sg.nchunks; ++ch, ++ia) {
                            const uint32_t sa = ia & 1;
                            ptx::mbar_wait(&afull_bar[sa], (ia >> 1) & 1);
                            ptx::tc_fence_after();
                            const uint32_t a0 = ptx::smem_u32(smem + sa * p.halo_a_stage) + w0 * 128;
                            for (int tap = 0; tap < sg.ntaps; ++tap, ++ib) {
                                const int r = (sg.ntaps == 9) ? tap / 3 : 1;
                                const int q = (sg.ntaps == 9) ? tap - 3 * r : 1;
                                const uint32_t sb = ib % SB;
                                ptx::mbar_wait(&full_bar[sb], (ib / SB) & 1);
                                ptx::tc_fence_after();
                                // shifted window of the halo tile: the swizzle is a function of the absolute smem address,
                                // so a start offset of any multiple of 128 bytes addresses the rows TMA wrote
                                // (verified on hardware, tools/exp_halo.py)
                                const uint64_t da = ptx::make_kmajor_sw128_desc(a0 + (r * Wp + q) * 128);
                                const uint64_t db = ptx::make_kmajor_sw128_desc(sB0 + sb * Cfg::B_STAGE_BYTES);
                                if (p.dbg_mode != 1) {
#pragma unroll
                                    for (int j = 0; j < TILE_K / 16; ++j) ptx::umma_f16(tacc, da + 2 * j, db + 2 * j, idesc, (first && j == 0) ? 0u : 1u);
                                }
                                first = 0;
                                ptx::umma_commit(&empty_bar[sb]);
                            }
                            ptx::umma_commit(&aempty_bar[sa]);
                        }
                    }
                    ptx::umma_commit(&tmem_full[acc]);
                }
            }
        } else if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(/*bf16*/ 1, TILE_M, BLOCK_N);
            uint32_t it = 0, ti = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
                const uint32_t acc = ti & 1;
                ptx::mbar_wait(&tmem_empty[acc], ((ti >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t tacc = tmem_base + acc * Cfg::ACC_COLS;
                for (int k = 0; k < k_iters; ++k, ++it) {
                    const uint32_t stage = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ptx::mbar_wait(&full_bar[stage], ph);
                    ptx::tc_fence_after();
                    if (it == 0) { DBG2(2); }
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(sa + A_STAGE_BYTES);
                    if (p.dbg_mode != 1) {
#pragma unroll
                        for (int j = 0; j < TILE_K / 16; ++j)
                            ptx::umma_f16(tacc, da + 2 * j, db + 2 * j, idesc, (k > 0 || j > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                }
                ptx::umma_commit(&tmem_full[acc]);
                if (ti == 0) { DBG2(3); }
                if (ti == 1) { DBG2(4); }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (8 warps, 256 threads): see epi_tile
        EpiCtx cx;
        cx.e = threadIdx.x - 64;
        cx.ew = cx.e >> 5;
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
        cx.hsel = cx.ew >> 2;
        const int row_in_tile = quarter * 32 + lane;
        cx.sw = row_in_tile & 7;
        cx.stage_off = (row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128;
        cx.slot0 = smem + Cfg::SM_OUT;
        cx.sst = reinterpret_cast<float2*>(smem + Cfg::SM_STAT);
        cx.taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        cx.rt0 = (cx.ew & 3) * 32 + (cx.ew >> 2) * 16 + (lane >> 3);  // warp -> 16 rows of one 32-row block
        cx.bu = lane & 7;
        cx.brs0 = (lane >> 3) == 0;
        const bool leader = (cx.e == 0);
        // launch-uniform epilogue shape
        const bool simple = !p.softmax && p.act == ACT_NONE && !p.out_fp32 && !(p.bias && p.bias_along_m);
        int mode = EPI_GENERIC;
        if (simple && !p.residual && !p.rowvec) mode = EPI_BIAS;
        else if (simple && !p.residual && p.rowvec) mode = EPI_ROWVEC;
        else if (simple && p.residual && !p.rowvec) mode = EPI_RESIDUAL;
        const bool has_stats = p.stats != nullptr && !p.out_fp32;
        uint32_t ti = 0, out_cnt = 0;

        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
            const int n_tile = t % p.n_tiles;
            const int mt = t / p.n_tiles;
            const int m_tile = mt % p.m_tiles;
            const int batch = mt / p.m_tiles;
            const int col0 = n_tile * BLOCK_N;
            int ncols = p.N_total - col0;
            if (ncols > BLOCK_N) ncols = BLOCK_N;
            const int nch = (ncols + 31) / 32;
            const uint32_t acc = ti & 1;

            ptx::mbar_wait(&tmem_full[acc], (ti >> 1) & 1);
            ptx::tc_fence_after();
            if (ti == 0 && leader) { DBG2(5); }
            const uint32_t tcol = acc * Cfg::ACC_COLS;
            uint64_t* te = &tmem_empty[acc];
            if (has_stats) {
                switch (mode) {
                    case EPI_BIAS: epi_tile<EPI_BIAS, true>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                    case EPI_ROWVEC: epi_tile<EPI_ROWVEC, true>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                    case EPI_RESIDUAL: epi_tile<EPI_RESIDUAL, true>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                    default: epi_tile<EPI_GENERIC, true>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                }
            } else {
                switch (mode) {
                    case EPI_BIAS: epi_tile<EPI_BIAS, false>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                    case EPI_ROWVEC: epi_tile<EPI_ROWVEC, false>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                    case EPI_RESIDUAL: epi_tile<EPI_RESIDUAL, false>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                    default: epi_tile<EPI_GENERIC, false>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt); break;
                }
            }
            if (ti == 0 && leader) { DBG2(6); }
            if (nch == 0) {  // cannot happen (n tiles are clipped on the host), but never leave the MMA warp waiting
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tmem_empty[acc]);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) { DBG2(7); }
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side

void gemm_set_error(const char* msg);

int conv_gemm_v2_ring_bytes(int block_n) {
    switch (block_n) {
        case 32: return Cfg2<32>::RING_BYTES;
        case 64: return Cfg2<64>::RING_BYTES;
        case 128: return Cfg2<128>::RING_BYTES;
        case 192: return Cfg2<192>::RING_BYTES;
        case 256: return Cfg2<256>::RING_BYTES;
        default: return 0;
    }
}

bool conv_gemm_v2_supported(const ConvGemmParams& p, int block_n) {
    if (block_n != 32 && block_n != 64 && block_n != 128 && block_n != 192 && block_n != 256) return false;
    const int eb = p.out_fp32 ? 4 : 2;
    if ((static_cast<long long>(p.ldo) * eb) % 16 || (reinterpret_cast<uintptr_t>(p.out) & 15)) return false;
    if (p.out_batch_stride && (p.out_batch_stride * eb) % 16) return false;
    if (p.N_total % 8 || p.out_nchw) return false;
    if (p.residual && ((static_cast<long long>(p.ldr) * 2) % 16 || (reinterpret_cast<uintptr_t>(p.residual) & 15) ||
                       (p.res_batch_stride * 2) % 16))
        return false;
    if (p.softmax && p.N_total != block_n) return false;
    return true;
}

template <int BLOCK_N>
static int launch2_t(const ConvGemmParams& p, cudaStream_t stream) {
    using Cfg = Cfg2<BLOCK_N>;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm2_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            gemm_set_error(cudaGetErrorString(e));
            return (int)e;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int total = p.m_tiles * p.n_tiles * p.batch_count;
    const int grid = total < num_sms ? total : num_sms;
    conv_gemm2_kernel<BLOCK_N><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gemm_set_error(cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int launch_conv_gemm_v2(const ConvGemmParams& p, int block_n, cudaStream_t stream) {
    switch (block_n) {
        case 32: return launch2_t<32>(p, stream);
        case 64: return launch2_t<64>(p, stream);
        case 128: return launch2_t<128>(p, stream);
        case 192: return launch2_t<192>(p, stream);
        case 256: return launch2_t<256>(p, stream);
        default: gemm_set_error("unsupported block_n"); return -4;
    }
}

}  // namespace dxmi
