// Persistent tcgen05 implicit-GEMM convolution kernel for CTA PAIRS (cta_group::2). Same operator contract as gemm_tc2.cu.
//
// Why: the single-CTA kernel is bound by the per-SM TMA ingest rate (~140 GB/s): a 128 x BLOCK_N tile needs 16 KB of A
// and BLOCK_N/4 KB of B per 64-deep K step.  Two CTAs of a cluster (two SMs of one TPC) compute a 256 x BLOCK_N tile
// together: each loads its own 128 A rows and only HALF of the B tile; one tcgen05.mma.cta_group::2 (M = 256, issued by the
// leader CTA) reads A from each CTA's shared memory and the two B halves from both, and writes 128 accumulator rows into
// each CTA's TMEM.  Per-SM ingest per K step drops from 16 + BLOCK_N/4 KB to 16 + BLOCK_N/8 KB.
// (Mechanics verified first in isolation on hardware: tools/csrc/exp_2cta.cu, tools/exp_2cta.py.)
//
// Barriers: full[s] lives in the LEADER (2 arrivals + the bytes of both CTAs; the peer's TMA loads signal it through the
// cluster window), empty[s] / tmem_full[a] live in BOTH CTAs (multicast tcgen05.commit), tmem_empty[a] lives in the leader
// (2 x 256 epilogue threads arrive, the peer's remotely).
#include "gemm_epi.cuh"

#include <cstdio>

namespace dxmi {

// warp group 0: TMA warp, MMA warp, two idle warps (56 registers each after setmaxnreg); warp groups 1-2: 8 epilogue warps (224)
static constexpr int NUM_THREADS_P = 384;

// RESB ("weights stationary"): when the whole B operand of this CTA (its BLOCK_N/2 rows x K) fits next to a shallow A
// ring, it is loaded ONCE per CTA and stays in shared memory for every tile the persistent CTA processes; the ring then
// carries only A.  The N=128 3x3 convs on 32x32 maps are bound by aggregate L2 -> SM bandwidth (14.5 TB/s measured: the
// same number for the one-CTA and the pair kernel, with the MMAs switched off), so removing B from the per-K-step traffic
// (24 KB -> 16 KB per SM) looked attractive.  MEASURED NEGATIVE (profiles/r01_pair_resident_b.txt): the 147 KB resident
// operand leaves room for only 3 A stages, and with 48 KB instead of 192 KB in flight per SM the loop becomes latency bound
// (32x32 128->128: 91 us vs 72 us streaming).  Off by default (option "pair_resident_b").
static constexpr int RESB_MAX_KITERS = 18;  // K <= 1152 (one 3x3 conv over 128 channels)
template <int BLOCK_N, bool RESB = false>
struct Cfg2P {
    static constexpr int B_STAGE_BYTES = (BLOCK_N / 2) * TILE_K * 2;  // this CTA's half of the B tile
    static constexpr int STAGE_BYTES = RESB ? A_STAGE_BYTES : A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = RESB ? 3 : (BLOCK_N > 128 ? 6 : 8);
    static constexpr int ACC_COLS = BLOCK_N <= 128 ? 128 : 256;
    static constexpr int TMEM_COLS = BLOCK_N == 128 ? 512 : 2 * ACC_COLS;  // (128: room for the two-tiles-per-CTA mode)
    static constexpr int SM_RESB = STAGES * STAGE_BYTES;  // resident B: RESB_MAX_KITERS boxes of B_STAGE_BYTES
    static constexpr int RING_BYTES = SM_RESB + (RESB ? RESB_MAX_KITERS * B_STAGE_BYTES : 0);
    static constexpr int SM_OUT = RING_BYTES;
    static constexpr int SM_STAT = SM_OUT + 2 * EPI_SLOT_BYTES;
    static constexpr int SM_BAR = SM_STAT + 2048;
    static constexpr int SMEM_BYTES = SM_BAR + 512;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(RING_BYTES % 1024 == 0, "staging slots must stay 1024-byte aligned");
};

__device__ __forceinline__ uint32_t ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

template <int BLOCK_N, bool RESB>
__global__ void __launch_bounds__(NUM_THREADS_P, 1) conv_gemm2p_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = Cfg2P<BLOCK_N, RESB>;
    constexpr int STAGES = Cfg::STAGES;
    const bool S3 = BLOCK_N == 128 && !RESB && p.shift3;  // launch-uniform
    // M2 (shift-3 only): a CTA owns two vertically adjacent tiles (256 rows) per step - one A box of 2*bh+2 image rows and ONE copy
    // of the B tiles feed both, which cuts the bytes the SM ingests per output tile by a third (these GEMMs are ingest bound)
    const int msub = (S3 && p.s3_m2) ? 2 : 1;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SM_BAR);
    uint64_t* full_bar = bars;                    // [STAGES]  (used in the leader)
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]       (used in the leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint64_t* resb_full = bars + 2 * STAGES + 5;  // RESB: the resident B operand of both CTAs has landed (leader's barrier)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;
    const int m_pairs = (p.m_tiles + 2 * msub - 1) / (2 * msub);
    const int total_pairs = m_pairs * p.n_tiles * p.batch_count;

    int k_iters = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s)
        if (s < p.nseg) k_iters += p.seg[s].ntaps * p.seg[s].nchunks;

    if (threadIdx.x == 0) {
        if (ptx::smem_u32(smem) & 1023u) {
            printf("dxmi conv_gemm2p: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tmap(&p.a_map[0]);
        if (p.nseg > 1) ptx::prefetch_tmap(&p.a_map[1]);
        if (p.nseg > 2) ptx::prefetch_tmap(&p.a_map[2]);
        ptx::prefetch_tmap(&p.b_map);
        if (S3) {
            ptx::prefetch_tmap(&p.s3_map[0]);
            if (p.nseg > 1) ptx::prefetch_tmap(&p.s3_map[1]);
            if (p.nseg > 2) ptx::prefetch_tmap(&p.s3_map[2]);
        }
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 2);   // one arrival per CTA (+ the transaction bytes of both)
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 512 * msub);  // the 256 epilogue threads of each CTA, once per tile they drain
        }
        ptx::mbar_init(resb_full, 2);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: the prologue above overlapped the previous kernel's tail; from here on its outputs are visible
    ptx::pdl_wait();
    ptx::pdl_trigger();

    if (warp < 4) {
    ptx::reg_dec<56>();
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (ptx::elect_one()) {
            const int tiles_per_nblk = p.tiles_w * p.tiles_h;
            uint32_t it = 0;
            if (RESB) {
                // one n tile, unbatched B: this CTA's half of every K slice, once
                const uint32_t rb_leader = mapa(ptx::smem_u32(resb_full), 0);
                if (rank == 0) {
                    ptx::mbar_expect_tx(resb_full, 2 * k_iters * Cfg::B_STAGE_BYTES);
                } else {
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rb_leader) : "memory");
                }
                // resident slot j holds the K slice the j-th stage of a tile consumes (chunk, column shift, row shift order)
                int j = 0, kb = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const GemmSeg sg = p.seg[s];
                    const int nq = sg.ntaps == 9 ? 3 : 1;
                    for (int ch = 0; ch < sg.nchunks; ++ch)
                        for (int q = 0; q < nq; ++q)
                            for (int r = 0; r < nq; ++r, ++j)
                                tma2_load_3d(smem + Cfg::SM_RESB + j * Cfg::B_STAGE_BYTES, &p.b_map, rb_leader,
                                             (kb + (r * 3 + q) * sg.nchunks + ch) * TILE_K, (int)rank * (BLOCK_N / 2), 0);
                    kb += sg.ntaps * sg.nchunks;
                }
            }
            for (int tp = cluster_id; tp < total_pairs; tp += n_clusters) {
                const int n_tile = tp % p.n_tiles;
                const int mp = tp / p.n_tiles;
                const int m_tile = ((mp % m_pairs) * 2 + (int)rank) * msub;
                const int batch = mp / m_pairs;
                const int n_blk = m_tile / tiles_per_nblk;
                const int rem = m_tile - n_blk * tiles_per_nblk;
                const int h_blk = rem / p.tiles_w;
                const int w_blk = rem - h_blk * p.tiles_w;
                const int w0 = w_blk * p.bw * p.stride;
                const int h0 = h_blk * p.bh * p.stride;
                const int n0 = p.a_batched ? batch : n_blk * p.bn;   // past the tensor for an odd last tile: zero fill
                const int bcoord_n = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);
                const int bcoord_b = p.b_batched ? batch : 0;
                int kk = 0;
                if (S3) {
                    // shift-3 ring: stage = [A box | 3 B tiles] for 3x3 segments, [A tile | B tile] for 1x1 segments
                    const int stage_bytes = p.s3_a_bytes + 3 * Cfg::B_STAGE_BYTES;
                    for (int s = 0; s < p.nseg; ++s) {
                        const GemmSeg sg = p.seg[s];
                        if (sg.ntaps == 9) {
                            for (int ch = 0; ch < sg.nchunks; ++ch) {
                                for (int q = 0; q < 3; ++q, ++it) {
                                    const uint32_t stage = it % p.s3_stages;
                                    const uint32_t ph = (it / p.s3_stages) & 1;
                                    ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
                                    uint8_t* sa = smem + stage * stage_bytes;
                                    uint8_t* sb = sa + p.s3_a_bytes;
                                    const uint32_t full_leader = mapa(ptx::smem_u32(&full_bar[stage]), 0);
                                    if (rank == 0) {
                                        ptx::mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);
                                    } else {
                                        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(full_leader) : "memory");
                                    }
                                    tma2_load_4d(sa, &p.s3_map[sg.map], full_leader, ch * TILE_K, q - 1, h0 - 1, n0);
#pragma unroll
                                    for (int r = 0; r < 3; ++r)
                                        tma2_load_3d(sb + r * Cfg::B_STAGE_BYTES, &p.b_map, full_leader,
                                                     (kk + (r * 3 + q) * sg.nchunks + ch) * TILE_K, bcoord_n, bcoord_b);
                                }
                            }
                            kk += 9 * sg.nchunks;
                        } else {
                            for (int ch = 0; ch < sg.nchunks; ++ch, ++it, ++kk) {
                                const uint32_t stage = it % p.s3_stages;
                                const uint32_t ph = (it / p.s3_stages) & 1;
                                ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
                                uint8_t* sa = smem + stage * stage_bytes;
                                uint8_t* sb = sa + p.s3_a_bytes;
                                const uint32_t full_leader = mapa(ptx::smem_u32(&full_bar[stage]), 0);
                                if (rank == 0) {
                                    ptx::mbar_expect_tx(&full_bar[stage], 2 * (msub * A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
                                } else {
                                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(full_leader) : "memory");
                                }
                                tma2_load_4d(sa, &p.a_map[sg.map], full_leader, ch * TILE_K, w0, h0, n0);
                                tma2_load_3d(sb, &p.b_map, full_leader, kk * TILE_K, bcoord_n, bcoord_b);
                            }
                        }
                    }
                    continue;
                }
                for (int s = 0; s < p.nseg; ++s) {
                    const GemmSeg sg = p.seg[s];
                    const CUtensorMap* amap = &p.a_map[sg.map];
                    // K order of a 3x3 segment: chunk, column shift q, row shift r (same as the shift-3 mode above and as gemm_tc2.cu)
                    const int nq = sg.ntaps == 9 ? 3 : (sg.ntaps == 4 ? 2 : 1);
                    // up2 mode: the batch index is the output phase (py, px); its 2x2 taps start at (py - 1, px - 1)
                    const int upx = p.up2 ? (batch & 1) : 0, upy = p.up2 ? (batch >> 1) : 0;
                    for (int ch = 0; ch < sg.nchunks; ++ch) {
                        for (int q = 0; q < nq; ++q) {
                            for (int r = 0; r < nq; ++r, ++it) {
                                const int kx = kk + (r * nq + q) * sg.nchunks + ch;
                                const uint32_t stage = it % STAGES;
                                const uint32_t ph = (it / STAGES) & 1;
                                ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
                                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                                uint8_t* sb = sa + A_STAGE_BYTES;
                                const uint32_t full_leader = mapa(ptx::smem_u32(&full_bar[stage]), 0);
                                if (rank == 0) {
                                    ptx::mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                                } else {
                                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(full_leader) : "memory");
                                }
                                tma2_load_4d(sa, amap, full_leader, ch * TILE_K, w0 + q - sg.pad + upx, h0 + r - sg.pad + upy, n0);
                                if (!RESB) tma2_load_3d(sb, &p.b_map, full_leader, kx * TILE_K, bcoord_n, bcoord_b);
                            }
                        }
                    }
                    kk += sg.ntaps * sg.nchunks;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(/*bf16*/ 1, 256, BLOCK_N);
            uint32_t it = 0, ti = 0;
            if (RESB) {
                ptx::mbar_wait(resb_full, 0);
                ptx::tc_fence_after();
            }
            for (int tp = cluster_id; tp < total_pairs; tp += n_clusters, ++ti) {
                const uint32_t acc = ti & 1;
                ptx::mbar_wait(&tmem_empty[acc], ((ti >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t tacc = tmem_base + acc * (Cfg::ACC_COLS * msub);
                if (S3) {
                    const int stage_bytes = p.s3_a_bytes + 3 * Cfg::B_STAGE_BYTES;
                    uint32_t first = 1;
                    for (int s = 0; s < p.nseg; ++s) {
                        const GemmSeg sg = p.seg[s];
                        const int nst = sg.ntaps == 9 ? 3 * sg.nchunks : sg.nchunks;
                        // smem offset of the second tile's rows inside the stage's A box / tile
                        const uint32_t sub_off = sg.ntaps == 9 ? (uint32_t)p.bh * p.s3_row_bytes : (uint32_t)A_STAGE_BYTES;
                        for (int k = 0; k < nst; ++k, ++it) {
                            const uint32_t stage = it % p.s3_stages;
                            const uint32_t ph = (it / p.s3_stages) & 1;
                            ptx::mbar_wait(&full_bar[stage], ph);
                            ptx::tc_fence_after();
                            const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
                            const uint32_t sb = sa + p.s3_a_bytes;
                            const int nr = sg.ntaps == 9 ? 3 : 1;
                            for (int r = 0; r < nr; ++r) {
                                const uint64_t db = ptx::make_kmajor_sw128_desc(sb + r * Cfg::B_STAGE_BYTES);
                                for (int h = 0; h < msub; ++h) {
                                    const uint64_t da = ptx::make_kmajor_sw128_desc(sa + r * p.s3_row_bytes + h * sub_off);
                                    if (p.dbg_mode != 1) {
#pragma unroll
                                        for (int j = 0; j < TILE_K / 16; ++j)
                                            umma2_f16(tacc + h * Cfg::ACC_COLS, da + 2 * j, db + 2 * j, idesc, (first && j == 0) ? 0u : 1u);
                                    }
                                }
                                first = 0;
                            }
                            umma2_commit_both(&empty_bar[stage]);
                        }
                    }
                    umma2_commit_both(&tmem_full[acc]);
                    continue;
                }
                for (int k = 0; k < k_iters; ++k, ++it) {
                    const uint32_t stage = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ptx::mbar_wait(&full_bar[stage], ph);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(
                        RESB ? ptx::smem_u32(smem + Cfg::SM_RESB + k * Cfg::B_STAGE_BYTES) : sa + A_STAGE_BYTES);
                    if (p.dbg_mode != 1) {
#pragma unroll
                        for (int j = 0; j < TILE_K / 16; ++j) umma2_f16(tacc, da + 2 * j, db + 2 * j, idesc, (k > 0 || j > 0) ? 1u : 0u);
                    }
                    umma2_commit_both(&empty_bar[stage]);
                }
                umma2_commit_both(&tmem_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ------------------------------------------------------------ GroupNorm statistics publisher: see epi_publish_tile
        if (epi_stats_published(p)) {
            const float* sstf = reinterpret_cast<const float*>(smem + Cfg::SM_STAT);
            ptx::named_bar_arrive(3, 288);
            for (int tp = cluster_id; tp < total_pairs; tp += n_clusters) {
                const int mp = tp / p.n_tiles;
                const int m_tile0 = ((mp % m_pairs) * 2 + (int)rank) * msub;
                const int col0 = (tp % p.n_tiles) * BLOCK_N;
                int ncols = p.N_total - col0;
                if (ncols > BLOCK_N) ncols = BLOCK_N;
                for (int h = 0; h < msub; ++h) epi_publish_tile(p, sstf, m_tile0 + h, col0, (ncols + 31) / 32, lane, mp / m_pairs);
            }
        }
    }
    } else {
        ptx::reg_inc<224>();
        // ------------------------------------------------------------ epilogue (8 warps per CTA): see gemm_epi.cuh
        EpiCtx cx;
        cx.e = threadIdx.x - 128;
        cx.ew = cx.e >> 5;
        const int quarter = warp & 3;
        cx.hsel = cx.ew >> 2;
        const int row_in_tile = quarter * 32 + lane;
        cx.sw = row_in_tile & 7;
        cx.stage_off = (row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128;
        cx.slot0 = smem + Cfg::SM_OUT;
        cx.sst = reinterpret_cast<float2*>(smem + Cfg::SM_STAT);
        cx.taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        cx.rt0 = (cx.ew & 3) * 32 + (cx.ew >> 2) * 16 + (lane >> 3);
        cx.bu = lane & 7;
        cx.brs0 = (lane >> 3) == 0;
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
            for (int tp = cluster_id; tp < total_pairs; tp += n_clusters, ++ti) {
                const int n_tile = tp % p.n_tiles;
                const int mp = tp / p.n_tiles;
                const int m_tile0 = ((mp % m_pairs) * 2 + (int)rank) * msub;
                const int batch = mp / m_pairs;
                const int col0 = n_tile * BLOCK_N;
                int ncols = p.N_total - col0;
                if (ncols > BLOCK_N) ncols = BLOCK_N;
                const int nch = (ncols + 31) / 32;
                const uint32_t acc = ti & 1;
                const uint32_t te = mapa(ptx::smem_u32(&tmem_empty[acc]), 0);  // the leader's barrier
                int nx_m = -1, nx_col0 = 0, nx_batch = 0;  // this CTA's next step (operand prefetch)
                if (tp + n_clusters < total_pairs) {
                    const int tn = tp + n_clusters;
                    const int mpn = tn / p.n_tiles;
                    nx_col0 = (tn % p.n_tiles) * BLOCK_N;
                    nx_m = ((mpn % m_pairs) * 2 + (int)rank) * msub;
                    nx_batch = mpn / m_pairs;
                }
                for (int h = 0; h < msub; ++h) {
                    const uint32_t tcol = acc * (Cfg::ACC_COLS * msub) + h * Cfg::ACC_COLS;
                    const bool last = h == msub - 1;
                    if (MODE == EPI_BIAS && !ST && epi_lean_ok(p))
                        epi_tile_lean(p, cx, tcol, te, m_tile0 + h, col0, nch, batch, &tmem_full[acc], (ti >> 1) & 1);
                    else
                        epi_tile<MODE, ST, RW>(p, cx, tcol, te, m_tile0 + h, col0, nch, batch, out_cnt, &tmem_full[acc], (ti >> 1) & 1, carry,
                                               last ? nx_m : m_tile0 + h + 1, last ? nx_col0 : col0, last ? nx_batch : batch);
                }
            }
#ifdef DXMI_EPI_PROFILE
            if (p.dbg_times && cx.e == 0)
                for (int k = 0; k < 8; ++k) p.dbg_times[(long long)blockIdx.x * 8 + k] = carry.prof[k];
#endif
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
    cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side

void gemm_set_error(const char* msg);

int conv_gemm_pair_ring_bytes(int block_n) {
    switch (block_n) {
        case 128: return Cfg2P<128, false>::RING_BYTES;
        case 192: return Cfg2P<192, false>::RING_BYTES;
        case 256: return Cfg2P<256, false>::RING_BYTES;
        default: return 0;
    }
}

bool conv_gemm_pair_supported(const ConvGemmParams& p, int block_n) {
    return (block_n == 128 || block_n == 192 || block_n == 256) && !p.halo && !p.softmax;
}

template <int BLOCK_N, bool RESB>
static int launch2p_t(const ConvGemmParams& p, cudaStream_t stream) {
    using Cfg = Cfg2P<BLOCK_N, RESB>;
    static DevFlags configured;
    static int num_sms = 0;
    if (!configured.test()) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm2p_kernel<BLOCK_N, RESB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            gemm_set_error(cudaGetErrorString(e));
            return (int)e;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.set();
    }
    const int msub = (BLOCK_N == 128 && !RESB && p.shift3 && p.s3_m2) ? 2 : 1;
    const int total_pairs = ((p.m_tiles + 2 * msub - 1) / (2 * msub)) * p.n_tiles * p.batch_count;
    int clusters = num_sms / 2;
    if (total_pairs < clusters) clusters = total_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS_P);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm2p_kernel<BLOCK_N, RESB>, p);
    if (e != cudaSuccess) {
        gemm_set_error(cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

static int g_opt_resb = 0;  // measured slower (see RESB comment): kept as an A/B switch
void set_pair_resident_b(int v) { g_opt_resb = v; }
int pair_resident_b_enabled() { return g_opt_resb; }

int launch_conv_gemm_pair(const ConvGemmParams& p, int block_n, cudaStream_t stream) {
    int k_iters = 0;
    for (int s = 0; s < p.nseg; ++s) k_iters += p.seg[s].ntaps * p.seg[s].nchunks;
    // weights-stationary variant: one column tile, unbatched B, K <= 1152, and enough tiles per CTA pair to amortise the preload
    if (g_opt_resb && block_n == 128 && p.n_tiles == 1 && !p.b_batched && k_iters <= RESB_MAX_KITERS &&
        ((p.m_tiles + 1) / 2) * p.batch_count >= 4 * 74)
        return launch2p_t<128, true>(p, stream);
    switch (block_n) {
        case 128: return launch2p_t<128, false>(p, stream);
        case 192: return launch2p_t<192, false>(p, stream);
        case 256: return launch2p_t<256, false>(p, stream);
        default: gemm_set_error("pair kernel: unsupported block_n"); return -4;
    }
}

}  // namespace dxmi
