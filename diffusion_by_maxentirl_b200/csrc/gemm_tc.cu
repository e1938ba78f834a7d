// tcgen05 implicit-GEMM convolution kernel. See gemm_tc.cuh for the operator contract.
//
// Warp roles (192 threads, one 128 x BLOCK_N output tile per CTA):
//   warp 0      : TMA producer (one elected lane) - A box + B box per K iteration into a STAGES-deep smem ring
//   warp 1      : TMEM allocator + MMA issuer (one elected lane) - 4 x tcgen05.mma (K=16) per K iteration
//   warps 2..5  : epilogue - tcgen05.ld 32 lanes x 32 columns, fused bias / temb / residual / activation /
//                 row-softmax, bf16 (or fp32) stores. Warp w owns TMEM lanes 32*(w%4) .. +31.
// Several CTAs are co-resident per SM (smem permitting) so one CTA's epilogue overlaps another's main loop.
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>

namespace dxmi {

static int g_opt_pdl = 1;
int pdl_enabled() { return g_opt_pdl; }
void set_pdl(int v) { g_opt_pdl = v; }


static constexpr int TILE_M = 128;
static constexpr int TILE_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
static constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
static constexpr int NUM_THREADS = 192;

template <int BLOCK_N>
struct TileCfg {
    static constexpr int B_STAGE_BYTES = BLOCK_N * TILE_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    // Keep >= 2 CTAs per SM for BLOCK_N <= 128 (3 x 32 KB = 96 KB) and a 4-deep ring for BLOCK_N = 256.
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 3 : 4);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
    static constexpr int TMEM_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;  // power of two
};

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DBG_STAMP(slot)                                                                                      \
    if (p.dbg_times) {                                                                                       \
        const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z); \
        p.dbg_times[cta * 8 + (slot)] = gtimer();                                                            \
    }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
    if (act == ACT_SILU) return v / (1.f + __expf(-v));
    return v;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = TileCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { DBG_STAMP(0); }

    const int m_tile = blockIdx.x;
    const int n_tile = blockIdx.y;
    const int batch = blockIdx.z;

    int k_iters = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s)
        if (s < p.nseg) k_iters += p.seg[s].ntaps * p.seg[s].nchunks;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.a_map[0]);
        if (p.nseg > 1) ptx::prefetch_tmap(&p.a_map[1]);
        if (p.nseg > 2) ptx::prefetch_tmap(&p.a_map[2]);
        ptx::prefetch_tmap(&p.b_map);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(&tmem_full_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (threadIdx.x == 0) { DBG_STAMP(1); }

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            const int tiles_per_nblk = p.tiles_w * p.tiles_h;
            const int n_blk = m_tile / tiles_per_nblk;
            const int rem = m_tile - n_blk * tiles_per_nblk;
            const int h_blk = rem / p.tiles_w;
            const int w_blk = rem - h_blk * p.tiles_w;
            const int w0 = w_blk * p.bw * p.stride;
            const int h0 = h_blk * p.bh * p.stride;
            const int n0 = p.a_batched ? batch : n_blk * p.bn;
            const int bcoord_n = n_tile * BLOCK_N;
            const int bcoord_b = p.b_batched ? batch : 0;

            int it = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const GemmSeg sg = p.seg[s];
                const CUtensorMap* amap = &p.a_map[sg.map];
                for (int tap = 0; tap < sg.ntaps; ++tap) {
                    const int r = (sg.ntaps == 9) ? tap / 3 : 0;
                    const int q = (sg.ntaps == 9) ? tap - 3 * r : 0;
                    for (int ch = 0; ch < sg.nchunks; ++ch, ++it) {
                        const int stage = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + A_STAGE_BYTES;
                        ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        ptx::tma_load_4d(sa, amap, &full_bar[stage], ch * TILE_K, w0 + q - sg.pad, h0 + r - sg.pad, n0);
                        ptx::tma_load_3d(sb, &p.b_map, &full_bar[stage], it * TILE_K, bcoord_n, bcoord_b);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(/*bf16*/ 1, TILE_M, BLOCK_N);
            for (int it = 0; it < k_iters; ++it) {
                const int stage = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                ptx::mbar_wait(&full_bar[stage], ph);
                ptx::tc_fence_after();
                if (it == 0) { DBG_STAMP(2); }
                const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint32_t sb = sa + A_STAGE_BYTES;
                const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                const uint64_t db = ptx::make_kmajor_sw128_desc(sb);
                if (p.dbg_mode != 1) {
#pragma unroll
                    for (int k = 0; k < TILE_K / 16; ++k) {
                        // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                        ptx::umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                }
                ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            }
            ptx::umma_commit(&tmem_full_bar);  // accumulator complete
            DBG_STAMP(3);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const long long row = static_cast<long long>(m_tile) * TILE_M + row_in_tile;  // row within this batch entry
        const bool row_ok = row < p.M_total;
        const int col0 = n_tile * BLOCK_N;

        ptx::mbar_wait(&tmem_full_bar, 0);
        ptx::tc_fence_after();
        if (threadIdx.x == 64) { DBG_STAMP(4); }

        const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float bias_m = (p.bias && p.bias_along_m && row_ok) ? p.bias[row] : 0.f;
        const float* rowvec = (p.rowvec && row_ok) ? p.rowvec + (row / p.rows_per_image) * p.ldrv : nullptr;
        const __nv_bfloat16* res =
            (p.residual && row_ok) ? p.residual + batch * p.res_batch_stride + row * p.ldr : nullptr;

        float sm_max = -INFINITY, sm_sum = 0.f;
        if (p.softmax) {
            // pass 1: online max / sum over the full row held in TMEM
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                uint32_t v[32];
                __syncwarp();
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
            sm_sum = 1.f / sm_sum;
        }

#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr_row + c, v);
            ptx::tmem_ld_wait();
            const int col = col0 + c;
            if (row_ok && col < p.N_total) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
            if (p.softmax) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __expf(f[j] - sm_max) * sm_sum;
            } else {
                if (p.bias) {
                    if (p.bias_along_m) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += bias_m;
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col + j < p.N_total) f[j] += __ldg(p.bias + col + j);
                    }
                }
                if (rowvec) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col + j < p.N_total) f[j] += __ldg(rowvec + col + j);
                }
                if (res) {
                    if (col + 32 <= p.N_total) {
                        const uint4* r4 = reinterpret_cast<const uint4*>(res + col);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 u = __ldg(r4 + q);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                f[q * 8 + 2 * e] += __uint_as_float(w[e] << 16);
                                f[q * 8 + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
                            }
                        }
                    } else {
                        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col + j < p.N_total) f[j] += __bfloat162float(res[col + j]);
                    }
                }
                if (p.act != ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
                }
            }
            if (p.out_nchw) {
                // lanes are consecutive pixels of one image: each column store is a coalesced 128-byte line
                const long long img = row / p.rows_per_image;
                const long long pix = row - img * p.rows_per_image;
                float* o = reinterpret_cast<float*>(p.out) + (img * p.N_total + col) * p.rows_per_image + pix;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col + j < p.N_total) o[(long long)j * p.rows_per_image] = f[j];
            } else if (p.out_fp32) {
                float* o = reinterpret_cast<float*>(p.out) + batch * p.out_batch_stride + row * p.ldo + col;
                if (col + 32 <= p.N_total && (p.ldo & 3) == 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        reinterpret_cast<float4*>(o)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                } else {
                    _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col + j < p.N_total) o[j] = f[j];
                }
            } else {
                __nv_bfloat16* o =
                    reinterpret_cast<__nv_bfloat16*>(p.out) + batch * p.out_batch_stride + row * p.ldo + col;
                if (col + 32 <= p.N_total && (p.ldo & 7) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(f[q * 8 + 0], f[q * 8 + 1]);
                        __nv_bfloat162 t1 = __floats2bfloat162_rn(f[q * 8 + 2], f[q * 8 + 3]);
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(f[q * 8 + 4], f[q * 8 + 5]);
                        __nv_bfloat162 t3 = __floats2bfloat162_rn(f[q * 8 + 6], f[q * 8 + 7]);
                        u.x = *reinterpret_cast<uint32_t*>(&t0);
                        u.y = *reinterpret_cast<uint32_t*>(&t1);
                        u.z = *reinterpret_cast<uint32_t*>(&t2);
                        u.w = *reinterpret_cast<uint32_t*>(&t3);
                        reinterpret_cast<uint4*>(o)[q] = u;
                    }
                } else {
                    _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col + j < p.N_total) o[j] = __float2bfloat16_rn(f[j]);
                }
            }
            }  // row_ok
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) { DBG_STAMP(5); }
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side

static thread_local char g_err[512] = "";
const char* gemm_last_error() { return g_err; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(sym);
    });
    return fn;
}

int make_act_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, long long w_stride, long long h_stride,
                 long long n_stride, int bw, int bh, int bn, int stride) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return -1;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)w_stride * 2, (cuuint64_t)h_stride * 2, (cuuint64_t)n_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)TILE_K, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof g_err,
                 "cuTensorMapEncodeTiled(act) failed: %d (C=%d W=%d H=%d N=%d ws=%lld hs=%lld ns=%lld box=%d,%d,%d s=%d)",
                 (int)r, C, W, H, N, w_stride, h_stride, n_stride, bw, bh, bn, stride);
        return -2;
    }
    return 0;
}

int make_mat_map(CUtensorMap* out, const void* base, int K, int rows, int batch, long long row_stride,
                 long long batch_stride, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return -1;
    }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)(batch > 1 ? batch_stride : row_stride * rows) * 2};
    cuuint32_t box[3] = {(cuuint32_t)TILE_K, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled(mat) failed: %d (K=%d rows=%d batch=%d rs=%lld bs=%lld box=%d)",
                 (int)r, K, rows, batch, row_stride, batch_stride, box_rows);
        return -2;
    }
    return 0;
}

int make_out_map(CUtensorMap* out, const void* base, int elem_bytes, int cols, int rows, int batch, long long row_stride,
                 long long batch_stride) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return -1;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * elem_bytes,
                             (cuuint64_t)(batch > 1 ? batch_stride : row_stride * rows) * elem_bytes};
    cuuint32_t box[3] = {(cuuint32_t)(128 / elem_bytes), 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled(out) failed: %d (cols=%d rows=%d batch=%d rs=%lld bs=%lld eb=%d)",
                 (int)r, cols, rows, batch, row_stride, batch_stride, elem_bytes);
        return -2;
    }
    return 0;
}

void gemm_set_error(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); }

template <int BLOCK_N>
static int launch_t(const ConvGemmParams& p, int m_tiles, int n_tiles, int batch, cudaStream_t stream) {
    using Cfg = TileCfg<BLOCK_N>;
    static DevFlags configured;
    if (!configured.test()) {
        cudaError_t e =
            cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured.set();
    }
    dim3 grid(m_tiles, n_tiles, batch);
    conv_gemm_kernel<BLOCK_N><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "conv_gemm launch: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int launch_conv_gemm(const ConvGemmParams& p, int block_n, int m_tiles, int n_tiles, int batch, cudaStream_t stream) {
    if (p.softmax && p.N_total != block_n) {
        snprintf(g_err, sizeof g_err, "softmax epilogue needs N_total == block_n (%d vs %d)", p.N_total, block_n);
        return -3;
    }
    switch (block_n) {
        case 32: return launch_t<32>(p, m_tiles, n_tiles, batch, stream);
        case 64: return launch_t<64>(p, m_tiles, n_tiles, batch, stream);
        case 128: return launch_t<128>(p, m_tiles, n_tiles, batch, stream);
        case 192: return launch_t<192>(p, m_tiles, n_tiles, batch, stream);
        case 256: return launch_t<256>(p, m_tiles, n_tiles, batch, stream);
        default:
            snprintf(g_err, sizeof g_err, "unsupported block_n %d", block_n);
            return -4;
    }
}

}  // namespace dxmi
