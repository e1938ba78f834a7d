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
#include "gemm_epi.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <cstdio>

namespace dxmi {

// warp group 0: TMA warp, MMA warp, two idle warps (56 registers each after setmaxnreg); warp groups 1-2: 8 epilogue warps (224)
static constexpr int NUM_THREADS = 384;

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
    // PDL: the prologue above overlapped the previous kernel's tail; from here on its outputs are visible
    ptx::pdl_wait();
    ptx::pdl_trigger();
    if (threadIdx.x == 0) { DBG2(0); DBG2(1); }

    if (warp < 4) {
    ptx::reg_dec<56>();
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
                int kbase = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const GemmSeg sg = p.seg[s];
                    const CUtensorMap* amap = &p.a_map[sg.map];
                    // K order of a 3x3 segment: channel chunk, then column shift q, then row shift r - the SAME order as the pair
                    // kernel's shift-3 mode, so every kernel variant accumulates identically (bitwise batch invariance: which
                    // variant runs depends on the batch size)
                    const int nq = sg.ntaps == 9 ? 3 : (sg.ntaps == 4 ? 2 : 1);
                    // up2 mode: the batch index is the output phase (py, px); its 2x2 taps start at (py - 1, px - 1)
                    const int upx = p.up2 ? (batch & 1) : 0, upy = p.up2 ? (batch >> 1) : 0;
                    for (int ch = 0; ch < sg.nchunks; ++ch) {
                        for (int q = 0; q < nq; ++q) {
                            for (int r = 0; r < nq; ++r, ++it) {
                                const int kk = kbase + (r * nq + q) * sg.nchunks + ch;
                                const uint32_t stage = it % STAGES;
                                const uint32_t ph = (it / STAGES) & 1;
                                ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
                                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                                uint8_t* sb = sa + A_STAGE_BYTES;
                                ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                                ptx::tma_load_4d(sa, amap, &full_bar[stage], ch * TILE_K, w0 + q - sg.pad + upx, h0 + r - sg.pad + upy, n0);
                                ptx::tma_load_3d(sb, &p.b_map, &full_bar[stage], kk * TILE_K, bcoord_n, bcoord_b);
                            }
                        }
                    }
                    kbase += sg.ntaps * sg.nchunks;
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
                                // (verified on hardware, tools/exp_halo.py + tools/csrc/exp_halo.cu)
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
    } else if (warp == 2) {
        // ------------------------------------------------------------ GroupNorm statistics publisher: see epi_publish_tile
        if (epi_stats_published(p)) {
            const float* sstf = reinterpret_cast<const float*>(smem + Cfg::SM_STAT);
            ptx::named_bar_arrive(3, 288);
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int m_tile = (t / p.n_tiles) % p.m_tiles;
                const int col0 = (t % p.n_tiles) * BLOCK_N;
                int ncols = p.N_total - col0;
                if (ncols > BLOCK_N) ncols = BLOCK_N;
                epi_publish_tile(p, sstf, m_tile, col0, (ncols + 31) / 32, lane, (t / p.n_tiles) / p.m_tiles);
            }
        }
    }
    } else {
        ptx::reg_inc<224>();
        // ------------------------------------------------------------ epilogue (8 warps, 256 threads): see epi_tile
        EpiCtx cx;
        cx.e = threadIdx.x - 128;
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
        const bool simple = !p.softmax && p.act == ACT_NONE && !p.out_fp32 && !(p.bias && p.bias_along_m) && !p.gate;
        int mode = EPI_GENERIC;
        if (simple && !p.residual && !p.rowvec) mode = EPI_BIAS;
        else if (simple && !p.residual && p.rowvec) mode = EPI_ROWVEC;
        else if (simple && p.residual && !p.rowvec) mode = EPI_RESIDUAL;
        const bool has_stats = p.stats != nullptr && !p.out_fp32;
        // one instantiation of the whole tile loop per launch-uniform epilogue shape: each carries only its own prefetch state
        auto run_tiles = [&](auto mode_c, auto stats_c) {
            constexpr int MODE = decltype(mode_c)::value;
            constexpr bool ST = decltype(stats_c)::value != 0;
            uint32_t ti = 0, out_cnt = 0;
            constexpr int RW = BLOCK_N == 192 ? 3 : (BLOCK_N == 64 ? 2 : 4);  // operand window (chunks): must divide BLOCK_N / 32
            EpiCarry<MODE, RW> carry;
            carry.tile_key = -1;
#ifdef DXMI_EPI_PROFILE
            for (int k = 0; k < 8; ++k) carry.prof[k] = 0;
#endif
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

                if (ti == 0 && leader) { DBG2(5); }
                const uint32_t tcol = acc * Cfg::ACC_COLS;
                const uint32_t te = ptx::smem_u32(&tmem_empty[acc]);  // own CTA: shared::cta addresses are valid shared::cluster ones
                int nx_m = -1, nx_col0 = 0, nx_batch = 0;  // this CTA's next tile (operand prefetch)
                if (t + (int)gridDim.x < total_tiles) {
                    const int tn = t + gridDim.x;
                    const int mtn = tn / p.n_tiles;
                    nx_col0 = (tn % p.n_tiles) * BLOCK_N;
                    nx_m = mtn % p.m_tiles;
                    nx_batch = mtn / p.m_tiles;
                }
                if (MODE == EPI_BIAS && !ST && epi_lean_ok(p))
                    epi_tile_lean(p, cx, tcol, te, m_tile, col0, nch, batch, &tmem_full[acc], (ti >> 1) & 1);
                else
                    epi_tile<MODE, ST, RW>(p, cx, tcol, te, m_tile, col0, nch, batch, out_cnt, &tmem_full[acc], (ti >> 1) & 1, carry, nx_m, nx_col0, nx_batch);
                if (ti == 0 && leader) { DBG2(6); }
                if (nch == 0) {  // cannot happen (n tiles are clipped on the host), but never leave the MMA warp waiting
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&tmem_empty[acc]);
                }
            }
        };
        if (has_stats) {
            switch (mode) {
                case EPI_BIAS: run_tiles(EpiSlot<EPI_BIAS>{}, EpiSlot<1>{}); break;
                case EPI_ROWVEC: run_tiles(EpiSlot<EPI_ROWVEC>{}, EpiSlot<1>{}); break;
                case EPI_RESIDUAL: run_tiles(EpiSlot<EPI_RESIDUAL>{}, EpiSlot<1>{}); break;
                default: run_tiles(EpiSlot<EPI_GENERIC>{}, EpiSlot<1>{}); break;
            }
        } else {
            switch (mode) {
                case EPI_BIAS: run_tiles(EpiSlot<EPI_BIAS>{}, EpiSlot<0>{}); break;
                case EPI_ROWVEC: run_tiles(EpiSlot<EPI_ROWVEC>{}, EpiSlot<0>{}); break;
                case EPI_RESIDUAL: run_tiles(EpiSlot<EPI_RESIDUAL>{}, EpiSlot<0>{}); break;
                default: run_tiles(EpiSlot<EPI_GENERIC>{}, EpiSlot<0>{}); break;
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
    static DevFlags configured;
    static int num_sms = 0;
    if (!configured.test()) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm2_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            gemm_set_error(cudaGetErrorString(e));
            return (int)e;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.set();
    }
    const int total = p.m_tiles * p.n_tiles * p.batch_count;
    const int grid = total < num_sms ? total : num_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm2_kernel<BLOCK_N>, p);
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
