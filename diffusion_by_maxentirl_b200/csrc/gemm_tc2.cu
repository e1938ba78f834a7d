// Persistent tcgen05 implicit-GEMM convolution kernel (v2). Same operator contract as gemm_tc.cu / gemm_tc.cuh.
//
// What changed against v1 (measured with per-CTA globaltimer stamps, profiles/r01_cta_timeline.txt: the row-per-thread
// direct-store epilogue cost 7.5 us per 128x256 tile against a 12 us main loop and nothing overlapped it):
//   * one CTA per SM, looping over output tiles (static round-robin), so barrier init / TMEM alloc / descriptor
//     prefetch are paid once per SM instead of once per tile;
//   * two TMEM accumulators: the MMA warp starts tile i+1 while the epilogue warps drain tile i;
//   * the epilogue stages 128-byte-wide column chunks in swizzled shared memory and writes them with TMA stores (full
//     128-byte lines, bounds clipped by the tensor map); residual tiles arrive the same way through TMA loads, one
//     chunk ahead;
//   * optional fused GroupNorm partial statistics: per 32-row segment and output column, (sum, sumsq) of the
//     bf16-rounded outputs via a warp reduce-scatter (31 shuffles per 32 columns), so the consumer GroupNorm needs no
//     statistics pass over HBM.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2..5 epilogue
// (warp w owns TMEM lanes 32*(w%4)..+31; lane 0 of warp 2 issues the epilogue's TMA loads / stores).
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <cstdio>

namespace dxmi {

static constexpr int TILE_M = 128;
static constexpr int TILE_K = 64;
static constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
static constexpr int NUM_THREADS = 192;
static constexpr int EPI_SLOT_BYTES = 128 * 128;  // 128 rows x 128 bytes

template <int BLOCK_N>
struct Cfg2 {
    static constexpr int B_STAGE_BYTES = BLOCK_N * TILE_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = BLOCK_N > 192 ? 3 : (BLOCK_N > 128 ? 4 : (BLOCK_N > 64 ? 4 : 6));
    static constexpr int ACC_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SM_OUT = RING_BYTES;                       // 2 output staging slots
    static constexpr int SM_RES = SM_OUT + 2 * EPI_SLOT_BYTES;      // 2 residual staging slots
    static constexpr int SM_BAR = SM_RES + 2 * EPI_SLOT_BYTES;      // mbarriers + TMEM slot
    static constexpr int SMEM_BYTES = SM_BAR + 256;
};

__device__ __forceinline__ float act2(float v, int act) {
    if (act == ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
    if (act == ACT_SILU) return v / (1.f + __expf(-v));
    return v;
}

// lane L ends with the sum over the warp of column L of v[0..31]
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int j = 0; j < w; ++j) {
            const float send = up ? v[j] : v[j + w];
            const float keep = up ? v[j + w] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0];
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_gemm2_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = Cfg2<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SM_BAR);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
    uint64_t* res_full = bars + 2 * STAGES + 4;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

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
        ptx::prefetch_tmap(&p.out_map);
        if (p.residual) ptx::prefetch_tmap(&p.res_map);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 128);
            ptx::mbar_init(&res_full[s], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
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
        if (ptx::elect_one()) {
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
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (thread <-> output row)
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const bool leader = (warp == 2 && lane == 0);
        const int sw = row_in_tile & 7;
        const int CH = p.out_fp32 ? 32 : 64;  // output columns per 128-byte staging row
        uint8_t* out_slot0 = smem + Cfg::SM_OUT;
        uint8_t* res_slot0 = smem + Cfg::SM_RES;
        const uint32_t row_off = (row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128;
        uint32_t ti = 0, out_cnt = 0, res_cnt = 0;

        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
            const int n_tile = t % p.n_tiles;
            const int mt = t / p.n_tiles;
            const int m_tile = mt % p.m_tiles;
            const int batch = mt / p.m_tiles;
            const int row0 = m_tile * TILE_M;
            const long long row = static_cast<long long>(row0) + row_in_tile;  // row within this batch entry
            const bool row_ok = row < p.M_total;
            const int col0 = n_tile * BLOCK_N;
            int ncols = p.N_total - col0;
            if (ncols > BLOCK_N) ncols = BLOCK_N;
            const int nch = (ncols + CH - 1) / CH;
            const uint32_t acc = ti & 1;

            if (p.residual && leader) {
                ptx::mbar_expect_tx(&res_full[res_cnt & 1], EPI_SLOT_BYTES);
                ptx::tma_load_3d(res_slot0 + (res_cnt & 1) * EPI_SLOT_BYTES, &p.res_map, &res_full[res_cnt & 1], col0, row0, batch);
            }
            ptx::mbar_wait(&tmem_full[acc], (ti >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t taddr_row = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
            const float bias_m = (p.bias && p.bias_along_m && row_ok) ? p.bias[row] : 0.f;
            const float* rowvec = (p.rowvec && row_ok) ? p.rowvec + (row / p.rows_per_image) * p.ldrv : nullptr;

            float sm_max = -INFINITY, sm_inv = 0.f;
            if (p.softmax) {
                float sm_sum = 0.f;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(taddr_row + c, v);
                    ptx::tmem_ld_wait();
                    float cmax = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) cmax = fmaxf(cmax, __uint_as_float(v[j]) * p.alpha);
                    const float nmax = fmaxf(sm_max, cmax);
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) part += __expf(__uint_as_float(v[j]) * p.alpha - nmax);
                    sm_sum = sm_sum * __expf(sm_max - nmax) + part;
                    sm_max = nmax;
                }
                sm_inv = 1.f / sm_sum;
            }

#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
                const int col = col0 + c * CH;
                const bool last = (c == nch - 1);
                if (p.residual && leader && !last) {
                    const uint32_t s = (res_cnt + 1) & 1;
                    ptx::mbar_expect_tx(&res_full[s], EPI_SLOT_BYTES);
                    ptx::tma_load_3d(res_slot0 + s * EPI_SLOT_BYTES, &p.res_map, &res_full[s], col + CH, row0, batch);
                }
                uint8_t* oslot = out_slot0 + (out_cnt & 1) * EPI_SLOT_BYTES + row_off;
                const uint8_t* rslot = res_slot0 + (res_cnt & 1) * EPI_SLOT_BYTES + row_off;
                if (p.residual) ptx::mbar_wait(&res_full[res_cnt & 1], (res_cnt >> 1) & 1);
                ptx::named_bar_sync(1, 128);  // staging slot (out_cnt & 1) has been read by its previous TMA store

                // two halves of 32 accumulator columns each (one half when the output is fp32)
                const int halves = p.out_fp32 ? 1 : 2;
                for (int hf = 0; hf < halves; ++hf) {
                    const int cc = col + hf * 32;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(taddr_row + (c * CH + hf * 32), v);
                    ptx::tmem_ld_wait();
                    if (last && hf == halves - 1) {
                        // every accumulator column of this tile is now in registers: hand the TMEM buffer back
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&tmem_empty[acc]);
                    }
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                    if (p.softmax) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __expf(f[j] - sm_max) * sm_inv;
                    } else {
                        if (p.bias) {
                            if (p.bias_along_m) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] += bias_m;
                            } else if (cc + 32 <= p.N_total) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cc) + q);
                                    f[4 * q] += b4.x;
                                    f[4 * q + 1] += b4.y;
                                    f[4 * q + 2] += b4.z;
                                    f[4 * q + 3] += b4.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (cc + j < p.N_total) f[j] += __ldg(p.bias + cc + j);
                            }
                        }
                        if (rowvec) {
                            if (cc + 32 <= p.N_total) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(rowvec + cc) + q);
                                    f[4 * q] += b4.x;
                                    f[4 * q + 1] += b4.y;
                                    f[4 * q + 2] += b4.z;
                                    f[4 * q + 3] += b4.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (cc + j < p.N_total) f[j] += __ldg(rowvec + cc + j);
                            }
                        }
                        if (p.residual) {
                            // residual slot: 128-byte rows of 64 bf16, 16-byte units XOR-swizzled by (row & 7)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int unit = (hf * 4 + q) ^ sw;
                                const uint4 u = *reinterpret_cast<const uint4*>(rslot + unit * 16);
                                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    f[q * 8 + 2 * e] += __uint_as_float(w[e] << 16);
                                    f[q * 8 + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
                                }
                            }
                        }
                        if (p.act != ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = act2(f[j], p.act);
                        }
                    }
                    if (p.out_fp32) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int unit = q ^ sw;
                            *reinterpret_cast<float4*>(oslot + unit * 16) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                        }
                    } else {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                            pk[j] = *reinterpret_cast<uint32_t*>(&b2);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int unit = (hf * 4 + q) ^ sw;
                            *reinterpret_cast<uint4*>(oslot + unit * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                        }
                        if (p.stats) {
                            // GroupNorm partials of the values the consumer will read (bf16-rounded), rows past M excluded
                            float s1[32], s2[32];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float a = row_ok ? __uint_as_float(pk[j] << 16) : 0.f;
                                const float b = row_ok ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
                                s1[2 * j] = a;
                                s1[2 * j + 1] = b;
                                s2[2 * j] = a * a;
                                s2[2 * j + 1] = b * b;
                            }
                            const float tsum = warp_reduce_scatter32(s1, lane);
                            const float tsq = warp_reduce_scatter32(s2, lane);
                            const long long seg = (static_cast<long long>(row0) + quarter * 32) >> 5;
                            if (cc + lane < p.N_total && static_cast<long long>(row0) + quarter * 32 < p.M_total)
                                *reinterpret_cast<float2*>(p.stats + (seg * p.N_total + cc + lane) * 2) = make_float2(tsum, tsq);
                        }
                    }
                }
                if (p.residual) ++res_cnt;
                ptx::fence_proxy_async_smem();
                ptx::named_bar_sync(2, 128);
                if (leader) {
                    if (p.dbg_mode != 2)
                        ptx::tma_store_3d(&p.out_map, out_slot0 + (out_cnt & 1) * EPI_SLOT_BYTES, col, row0, batch);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<1>();  // the store issued one chunk earlier has finished reading its slot
                }
                ++out_cnt;
            }
            if (nch == 0) {  // cannot happen (n tiles are clipped on the host), but never leave the MMA warp waiting
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tmem_empty[acc]);
            }
        }
        if (leader) ptx::bulk_wait<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side

void gemm_set_error(const char* msg);

bool conv_gemm_v2_supported(const ConvGemmParams& p, int block_n) {
    if (block_n != 32 && block_n != 64 && block_n != 128 && block_n != 192 && block_n != 256) return false;
    const int eb = p.out_fp32 ? 4 : 2;
    if ((static_cast<long long>(p.ldo) * eb) % 16 || (reinterpret_cast<uintptr_t>(p.out) & 15)) return false;
    if (p.out_batch_stride && (p.out_batch_stride * eb) % 16) return false;
    if (p.N_total % 8 || p.out_nchw) return false;
    if (p.residual && ((static_cast<long long>(p.ldr) * 2) % 16 || (reinterpret_cast<uintptr_t>(p.residual) & 15) || p.out_fp32))
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
