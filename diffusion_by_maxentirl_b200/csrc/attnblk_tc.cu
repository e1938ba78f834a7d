// Whole DDPM AttnBlock (models/DxMI/unet_small.py:167-191) at the 16x16 map (seq 256, C = 256, one head) as ONE kernel:
//
//     h = GroupNorm(x);  q, k, v = 1x1 convs of h;  w = softmax(q k^T C^-1/2);  o = w v;  out = x + proj_out(o)
//
// Before (round 1): GroupNorm finalize + apply, a q|k GEMM, a V^T GEMM, the fused S/softmax/PV kernel and a proj_out GEMM -
// five HBM round trips of [B,256,256] tensors and three K = 256 GEMM launches that were epilogue bound at 0.2-0.4 PFLOP/s
// (profiles/r01_gemm_table_cifar.txt: ~198 us per block at B = 256).  Here one 2-CTA cluster owns one image: x is read
// once, the six 256^3 contractions run back to back on tcgen05 with every intermediate (h, q, k, v^T, P, o) living in
// shared memory / TMEM, and only `out` (+ its GroupNorm partial statistics for the consumer) goes back to HBM.
//
// Mapping (cta_group::2, M = 256 = the image's 256 pixels, 128 per CTA; every operand K-major SWIZZLE_128B):
//     K   [T_A] = hn . Wk^T      A = hn   (own 128 rows)          B = Wk   (N split: this CTA streams rows rank*128..)
//     V^T [T_B] = Wv . hn^T      A = Wv   (M split: own 128 d)     B = hn   (N split = keys: own 128 rows)
//     Q   [T_A] = hn . Wq^T
//     S   [T_B] = Q . K^T        A = Q    (own queries)           B = K    (N split = keys: own 128 keys)
//     O   [T_A] = P . V          A = P    (own queries)           B = V^T  (N split = d: own 128 d, all 256 keys)
//     Y   [T_B] = O . Wp^T       A = O                              B = Wp   (N split)
// so the natural "own rows" placement of every intermediate is exactly what the pair MMA wants - no exchange of K / V between
// the two CTAs is ever needed.  Shared memory per CTA: three 64 KB operand regions (x/hn -> Q -> P -> out staging | K -> O | V^T)
// + a 2 x 16 KB weight ring + 2 KB of tables = 226 KB.  TMEM: two 256-column accumulators.
//
// Warp roles (320 threads): warps 0-7 drain / transform / softmax (warp w reads TMEM lanes 32*(w%4).., column half w/4),
// warp 8 = TMA producer, warp 9 = TMEM alloc + (leader CTA) MMA issuer.  Every barrier is one-shot except the weight ring.
#include "attn_tc.cuh"
#include "ptx.cuh"

#include <cstdio>

namespace dxmi {

static constexpr int AB_THREADS = 320;
static constexpr int AB_R0 = 0;
static constexpr int AB_R1 = 64 * 1024;
static constexpr int AB_R2 = 128 * 1024;
static constexpr int AB_WR = 192 * 1024;           // 2 stages x 16 KB
static constexpr int AB_TAB = 224 * 1024;          // float2[256] GroupNorm affine; later softmax row exchange
static constexpr int AB_BAR = AB_TAB + 2048;
static constexpr int AB_SMEM = AB_BAR + 256;
static_assert(AB_SMEM <= 227 * 1024, "shared memory budget");
static constexpr uint32_t AB_TA = 0, AB_TB = 256, AB_TCOLS = 512;

namespace {

__device__ __forceinline__ uint32_t ab_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void ab_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ab_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void ab_remote_arrive(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void ab_tma2_load_3d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void ab_umma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void ab_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ long long ab_gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define AB_STAMP(slot) \
    if (p.dbg) p.dbg[(long long)blockIdx.x * 16 + (slot)] = ab_gtimer();
__device__ __forceinline__ float ab_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// acquire-wait on a barrier the PEER's threads (or multicast commits) arrive on
__device__ __forceinline__ void ab_wait(uint64_t* bar, uint32_t parity) { ptx::mbar_wait(bar, parity); }

// bf16-pack 32 accumulator columns (value = acc * scale + bias_col[j] + bias_row) into 4 swizzled 16-byte units of the
// thread's row of a K-major SWIZZLE_128B panel (row r at r * 128 bytes, unit u stored at u ^ (r & 7)).
__device__ __forceinline__ void pack_store32(const uint32_t (&v)[32], float scale, const float* bias_col_s, float bias_row, uint8_t* dst_row,
                                             int u0, int sw) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (bias_col_s) {  // broadcast reads: every lane of the warp reads the same address
            b0 = *reinterpret_cast<const float4*>(bias_col_s + 8 * u);
            b1 = *reinterpret_cast<const float4*>(bias_col_s + 8 * u + 4);
        }
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a0 = fmaf(__uint_as_float(v[8 * u + 2 * j]), scale, bb[2 * j] + bias_row);
            const float a1 = fmaf(__uint_as_float(v[8 * u + 2 * j + 1]), scale, bb[2 * j + 1] + bias_row);
            __nv_bfloat162 t = __floats2bfloat162_rn(a0, a1);
            pk[j] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(dst_row + (((u0 + u) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// One thread's share of a TMEM -> shared-memory drain: its accumulator row (TMEM lane), 128 of the 256 columns, two
// 32-column TMEM loads in flight per iteration; `bias_col_s` (shared memory, 256 floats) may be null.
__device__ __forceinline__ void drain_half(uint32_t t_addr, uint8_t* region, int row, int half, float scale,
                                           const float* bias_col_s, float bias_row) {
    const int sw = row & 7;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int c0 = half * 128 + it * 64;  // one 64-column panel per iteration
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32b_x32(t_addr + c0, v0);
        ptx::tmem_ld_32x32b_x32(t_addr + c0 + 32, v1);
        ptx::tmem_ld_wait();
        uint8_t* dst = region + (c0 >> 6) * 16384 + row * 128;
        pack_store32(v0, scale, bias_col_s ? bias_col_s + c0 : nullptr, bias_row, dst, 0, sw);
        pack_store32(v1, scale, bias_col_s ? bias_col_s + c0 + 32 : nullptr, bias_row, dst, 4, sw);
    }
}

}  // namespace

__global__ void __launch_bounds__(AB_THREADS, 1) attnblk256_kernel(const __grid_constant__ AttnBlkParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AB_BAR);
    uint64_t* x_full = bars + 0;     // own: the CTA's 128 x 256 slice of x landed
    uint64_t* hn_ready = bars + 1;   // leader: both CTAs normalised their slice (count 2)
    uint64_t* w_full = bars + 2;     // [2] leader: weight stage landed in both CTAs
    uint64_t* w_empty = bars + 4;    // [2] both: stage consumed (multicast commit)
    uint64_t* c_done = bars + 6;     // [6] both: K, V^T, Q, S, O, Y accumulators complete (multicast commit)
    uint64_t* d_done = bars + 12;    // [5] leader: K, V^T, Q, P, O drained into shared memory by both CTAs (count 2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ab_ctarank();
    const int img = blockIdx.x >> 1;

    if (threadIdx.x == 0) {
        if (ptx::smem_u32(smem) & 1023u) {
            printf("dxmi attnblk256: dynamic smem base not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tmap(&p.x_map);
        ptx::prefetch_tmap(&p.w_map);
        ptx::mbar_init(x_full, 1);
        ptx::mbar_init(hn_ready, 2);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&w_full[s], 2);
            ptx::mbar_init(&w_empty[s], 1);
        }
        for (int i = 0; i < 6; ++i) ptx::mbar_init(&c_done[i], 1);
        for (int i = 0; i < 5; ++i) ptx::mbar_init(&d_done[i], 2);
        ptx::fence_mbar_init();
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                     "r"(AB_TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    ab_cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    ptx::pdl_wait();  // PDL: the prologue overlapped the previous kernel's tail
    ptx::pdl_trigger();

    if (warp == 8) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(x_full, 64 * 1024);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                ptx::tma_load_3d(smem + AB_R0 + c * 16384, &p.x_map, x_full, c * 64, (int)rank * 128, img);
            // weights: k, v, q, proj_out stacked as [1024][256]; this CTA streams rows mat*256 + rank*128 .. +127
            for (int i = 0; i < 16; ++i) {
                const int s = i & 1;
                const uint32_t k = i >> 1;
                ptx::mbar_wait(&w_empty[s], (k & 1) ^ 1);
                const uint32_t full_leader = ab_mapa(ptx::smem_u32(&w_full[s]), 0);
                if (rank == 0) {
                    ptx::mbar_expect_tx(&w_full[s], 2 * 16384);
                } else {
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(full_leader) : "memory");
                }
                ab_tma2_load_3d(smem + AB_WR + s * 16384, &p.w_map, full_leader, (i & 3) * 64, (i >> 2) * 256 + (int)rank * 128, 0);
            }
            // the residual: x again (L2 hit), into R2 once O = P V has consumed V^T - same swizzled panels as the output staging
            ptx::mbar_wait(&c_done[4], 0);
            ptx::mbar_expect_tx(x_full, 64 * 1024);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                ptx::tma_load_3d(smem + AB_R2 + c * 16384, &p.x_map, x_full, c * 64, (int)rank * 128, img);
        }
        __syncwarp();
    } else if (warp == 9) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
        if (rank == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(/*bf16*/ 1, 256, 256);
            const uint32_t r0 = ptx::smem_u32(smem + AB_R0), r1 = ptx::smem_u32(smem + AB_R1), r2 = ptx::smem_u32(smem + AB_R2);
            const uint32_t wr = ptx::smem_u32(smem + AB_WR);
            uint32_t wi = 0;  // weight stage counter
            auto gemm_w = [&](uint32_t tacc, uint32_t act_region, bool w_is_a) {
                for (int c = 0; c < 4; ++c, ++wi) {
                    const uint32_t s = wi & 1;
                    ptx::mbar_wait(&w_full[s], (wi >> 1) & 1);
                    ptx::tc_fence_after();
                    const uint64_t dact = ptx::make_kmajor_sw128_desc(act_region + c * 16384);
                    const uint64_t dw = ptx::make_kmajor_sw128_desc(wr + s * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ab_umma2(tacc, (w_is_a ? dw : dact) + 2 * k, (w_is_a ? dact : dw) + 2 * k, idesc, (c | k) ? 1u : 0u);
                    ab_commit_both(&w_empty[s]);
                }
            };
            auto gemm_ss = [&](uint32_t tacc, uint32_t a_region, uint32_t b_region) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint64_t da = ptx::make_kmajor_sw128_desc(a_region + c * 16384);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(b_region + c * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ab_umma2(tacc, da + 2 * k, db + 2 * k, idesc, (c | k) ? 1u : 0u);
                }
            };
            ab_wait(hn_ready, 0);
            ptx::tc_fence_after();
            AB_STAMP(8);
            gemm_w(tmem + AB_TA, r0, false);  // K = hn Wk^T
            ab_commit_both(&c_done[0]);
            gemm_w(tmem + AB_TB, r0, true);   // V^T = Wv hn^T
            ab_commit_both(&c_done[1]);
            ab_wait(&d_done[0], 0);           // K drained: T_A free
            ptx::tc_fence_after();
            AB_STAMP(9);
            gemm_w(tmem + AB_TA, r0, false);  // Q = hn Wq^T
            ab_commit_both(&c_done[2]);
            AB_STAMP(10);
            ab_wait(&d_done[1], 0);           // V^T drained: T_B free
            ab_wait(&d_done[2], 0);           // Q in shared memory (both CTAs)
            ptx::tc_fence_after();
            AB_STAMP(11);
            gemm_ss(tmem + AB_TB, r0, r1);    // S = Q K^T
            ab_commit_both(&c_done[3]);
            ab_wait(&d_done[3], 0);           // P written, S consumed
            ptx::tc_fence_after();
            AB_STAMP(12);
            gemm_ss(tmem + AB_TA, r0, r2);    // O = P V
            ab_commit_both(&c_done[4]);
            ab_wait(&d_done[4], 0);           // O in shared memory
            ptx::tc_fence_after();
            AB_STAMP(13);
            gemm_w(tmem + AB_TB, r1, false);  // Y = O Wp^T
            ab_commit_both(&c_done[5]);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ 8 drain warps (256 threads)
        const int tid = threadIdx.x;            // 0..255
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;    // accumulator row (TMEM lane) == pixel rank*128 + row
        const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t leader_d = ab_mapa(ptx::smem_u32(d_done), 0);
        auto publish = [&](int which) {
            // this CTA's share of a drained operand is in shared memory, its TMEM reads are complete
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1, 256);
            if (tid == 0) ab_remote_arrive(leader_d + which * 8);
        };

        float* tabf = reinterpret_cast<float*>(smem + AB_TAB);  // 512 floats: bias_k | bias_q, later softmax exchange, later bias_p
        tabf[tid] = __ldg(p.bias + tid);
        tabf[256 + tid] = __ldg(p.bias + 512 + tid);
        const float bias_v_row = __ldg(p.bias + 256 + rank * 128 + row);

        // ---- GroupNorm(32 groups of 8 channels): this thread owns channel unit cu (= group cu) of rows rg, rg + 8, ...
        const int cu = tid & 31, rg = tid >> 5;
        float ga[8], gb[8];
        {
            float s = 0.f, q = 0.f;
            const float4* base = reinterpret_cast<const float4*>(p.stats_in + (static_cast<long long>(img) * p.P_in * 256 + cu * 8) * 2);
            for (int seg = 0; seg < p.P_in; ++seg) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 v = __ldg(base + static_cast<long long>(seg) * 128 + e);  // (sum, sumsq) of two channels
                    s += v.x + v.z;
                    q += v.y + v.w;
                }
            }
            const float cnt = 8.f * 256.f;
            const float mean = s / cnt;
            const float var = fmaxf(q / cnt - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.eps);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + cu * 8)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + cu * 8) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + cu * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + cu * 8) + 1);
            const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                ga[e] = rstd * gm[e];
                gb[e] = bt[e] - mean * ga[e];
            }
        }
        if (tid == 0) { AB_STAMP(0); }
        ptx::mbar_wait(x_full, 0);
        if (tid == 0) { AB_STAMP(1); }
        // ---- hn = a * x + b in place: a warp touches one 512-byte row slice set per step (conflict free)
        {
            uint8_t* pbase = smem + AB_R0 + (cu >> 3) * 16384;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = rg + 8 * i;
                uint4* ptr = reinterpret_cast<uint4*>(pbase + r * 128 + (((cu & 7) ^ (r & 7)) << 4));
                uint4 v = *ptr;
                uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x0 = __uint_as_float(w[e] << 16), x1 = __uint_as_float(w[e] & 0xffff0000u);
                    __nv_bfloat162 t = __floats2bfloat162_rn(fmaf(x0, ga[2 * e], gb[2 * e]), fmaf(x1, ga[2 * e + 1], gb[2 * e + 1]));
                    w[e] = *reinterpret_cast<uint32_t*>(&t);
                }
                *ptr = v;
            }
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 256);
        if (tid == 0) ab_remote_arrive(ab_mapa(ptx::smem_u32(hn_ready), 0));

        // ---- K -> R1 (+ bias_k), V^T -> R2 (+ bias_v per row = output channel), Q -> R0 (+ bias_q)
        if (tid == 0) { AB_STAMP(2); }
        ptx::mbar_wait(&c_done[0], 0);
        ptx::tc_fence_after();
        if (tid == 0) { AB_STAMP(3); }
        drain_half(t_row + AB_TA, smem + AB_R1, row, half, 1.f, tabf, 0.f);
        publish(0);
        ptx::mbar_wait(&c_done[1], 0);
        ptx::tc_fence_after();
        drain_half(t_row + AB_TB, smem + AB_R2, row, half, 1.f, nullptr, bias_v_row);
        publish(1);
        ptx::mbar_wait(&c_done[2], 0);
        ptx::tc_fence_after();
        drain_half(t_row + AB_TA, smem + AB_R0, row, half, 1.f, tabf + 256, 0.f);
        publish(2);  // (its barrier also retires every read of the bias table)

        // ---- softmax over the 256 keys of this thread's query row (this thread: keys half*128 .. +127)
        float* xmax = tabf;        // [128][2]
        float* xsum = tabf + 256;  // [128][2]
        if (tid == 0) { AB_STAMP(4); }
        ptx::mbar_wait(&c_done[3], 0);
        ptx::tc_fence_after();
        if (tid == 0) { AB_STAMP(5); }
        float mx = -INFINITY;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            uint32_t v0[32], v1[32];
            ptx::tmem_ld_32x32b_x32(t_row + AB_TB + half * 128 + it * 64, v0);
            ptx::tmem_ld_32x32b_x32(t_row + AB_TB + half * 128 + it * 64 + 32, v1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
        }
        xmax[row * 2 + half] = mx;
        ptx::named_bar_sync(1, 256);
        mx = fmaxf(xmax[row * 2], xmax[row * 2 + 1]);
        const float m2 = mx * p.scale_log2;
        float rs = 0.f;
        {
            const int sw = row & 7;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int c0 = half * 128 + it * 64;
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32b_x32(t_row + AB_TB + c0, v0);
                ptx::tmem_ld_32x32b_x32(t_row + AB_TB + c0 + 32, v1);
                ptx::tmem_ld_wait();
                uint8_t* dst = smem + AB_R0 + (c0 >> 6) * 16384 + row * 128;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const uint32_t(&v)[32] = h2 ? v1 : v0;
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ab_exp2(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m2));
                        const float p1 = ab_exp2(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m2));
                        rs += p0 + p1;
                        __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
                        pk[j] = *reinterpret_cast<uint32_t*>(&t);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(dst + (((h2 * 4 + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
        }
        xsum[row * 2 + half] = rs;
        publish(3);  // (its named barrier also orders the xsum exchange and retires every read of xmax)
        const float inv = 1.f / (xsum[row * 2] + xsum[row * 2 + 1]);
        tabf[tid] = __ldg(p.bias + 768 + tid);  // proj_out bias into the (dead) xmax half; visible after publish(4)'s barrier

        // ---- O / rowsum -> R1
        ptx::mbar_wait(&c_done[4], 0);
        ptx::tc_fence_after();
        drain_half(t_row + AB_TA, smem + AB_R1, row, half, inv, nullptr, 0.f);
        publish(4);
        if (tid == 0) { AB_STAMP(6); }

        // ---- Y + bias_p + x (TMA-reloaded into R2) -> bf16 -> R0 staging
        ptx::mbar_wait(x_full, 1);
        ptx::mbar_wait(&c_done[5], 0);
        ptx::tc_fence_after();
        if (tid == 0) { AB_STAMP(7); }
        {
            const int sw = row & 7;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int c0 = half * 128 + it * 64;
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32b_x32(t_row + AB_TB + c0, v0);
                ptx::tmem_ld_32x32b_x32(t_row + AB_TB + c0 + 32, v1);
                const int poff = (c0 >> 6) * 16384 + row * 128;
                uint4 xr[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) xr[u] = *reinterpret_cast<const uint4*>(smem + AB_R2 + poff + ((u ^ sw) << 4));
                ptx::tmem_ld_wait();
                uint8_t* dst = smem + AB_R0 + poff;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t(&v)[32] = (u < 4) ? v0 : v1;
                    const float4 b0 = *reinterpret_cast<const float4*>(tabf + c0 + 8 * u);
                    const float4 b1 = *reinterpret_cast<const float4*>(tabf + c0 + 8 * u + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xr[u]);
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float a0 = __uint_as_float(v[8 * (u & 3) + 2 * j]) + bb[2 * j] + __uint_as_float(xw[j] << 16);
                        const float a1 = __uint_as_float(v[8 * (u & 3) + 2 * j + 1]) + bb[2 * j + 1] + __uint_as_float(xw[j] & 0xffff0000u);
                        __nv_bfloat162 t = __floats2bfloat162_rn(a0, a1);
                        pk[j] = *reinterpret_cast<uint32_t*>(&t);
                    }
                    *reinterpret_cast<uint4*>(dst + ((u ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        ptx::named_bar_sync(1, 256);
        // ---- coalesced stores (a warp writes one full 512-byte row) + GroupNorm partials of the bf16 outputs
        {
            const int unit = cu;  // 8 channels; rows rg, rg + 8, ...
            float s1[8], s2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
            __nv_bfloat16* obase = p.out + (static_cast<long long>(img) * 256 + rank * 128) * 256 + unit * 8;
            const uint8_t* sbase = smem + AB_R0 + (unit >> 3) * 16384;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = rg + 8 * i;
                const uint4 val = *reinterpret_cast<const uint4*>(sbase + r * 128 + (((unit & 7) ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(obase + static_cast<long long>(r) * 256) = val;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&val);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a0 = __uint_as_float(w[e] << 16), a1 = __uint_as_float(w[e] & 0xffff0000u);
                    s1[2 * e] += a0;
                    s2[2 * e] = fmaf(a0, a0, s2[2 * e]);
                    s1[2 * e + 1] += a1;
                    s2[2 * e + 1] = fmaf(a1, a1, s2[2 * e + 1]);
                }
            }
            if (p.stats_out) {
                // fixed-order combine of the 8 row groups through shared memory (R1 is free: O was consumed by Y = O Wp^T)
                float2* red = reinterpret_cast<float2*>(smem + AB_R1);  // [8][256]
#pragma unroll
                for (int e = 0; e < 8; ++e) red[rg * 256 + unit * 8 + e] = make_float2(s1[e], s2[e]);
                ptx::named_bar_sync(1, 256);
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float2 v = red[g * 256 + tid];
                    acc.x += v.x;
                    acc.y += v.y;
                }
                *reinterpret_cast<float2*>(p.stats_out + ((static_cast<long long>(img) * 2 + rank) * 256 + tid) * 2) = acc;
            }
        }
    }

    if (threadIdx.x == 0) { AB_STAMP(14); }
    ptx::tc_fence_before();
    ab_cluster_sync();
    if (warp == 9) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AB_TCOLS) : "memory");
    }
}

static long long* g_ab_dbg = nullptr;
void set_attnblk_dbg(void* p) { g_ab_dbg = (long long*)p; }
long long* attnblk_dbg_buffer() { return g_ab_dbg; }

int prepare_attnblk256(const void* x, const void* w_kvqp, const float* bias_kvqp, const float* gamma, const float* beta,
                       const float* stats_in, int P_in, float eps, float scale, void* out, float* stats_out, int B, AttnBlkOp* op) {
    AttnBlkParams& p = op->p;
    int r = make_mat_map(&p.x_map, x, 256, 256, B, 256, 256LL * 256, 128);
    if (!r) r = make_mat_map(&p.w_map, w_kvqp, 256, 1024, 1, 256, 0, 128);
    if (r) return r;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.stats_in = stats_in;
    p.P_in = P_in;
    p.stats_out = stats_out;
    p.gamma = gamma;
    p.beta = beta;
    p.bias = bias_kvqp;
    p.eps = eps;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.dbg = attnblk_dbg_buffer();
    op->B = B;
    op->flops = 6.0 * 2.0 * 256.0 * 256.0 * 256.0 * B;
    return 0;
}

int run_attnblk256(const AttnBlkOp& op, cudaStream_t st) {
    static DevFlags configured;
    if (!configured.test()) {
        cudaError_t e = cudaFuncSetAttribute(attnblk256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * op.B);
    cfg.blockDim = dim3(AB_THREADS);
    cfg.dynamicSmemBytes = AB_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, attnblk256_kernel, op.p);
    return (int)e;
}

}  // namespace dxmi
