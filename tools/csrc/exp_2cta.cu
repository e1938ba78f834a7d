// EXPERIMENT (not on the product path): the cta_group::2 mechanics - paired TMEM allocation, TMA loads that signal the
// leader CTA's mbarrier, one tcgen05.mma.cta_group::2 (M = 256 across the pair, B split along N between the two CTAs'
// shared memories), multicast commit - on a plain GEMM  D[256 x 128] = A[256 x K] . B[128 x K]^T  per cluster.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace dxmi {

struct Exp2Params {
    CUtensorMap a_map;  // (k, rows, 1) box (64, 128)
    CUtensorMap b_map;  // (k, rows, 1) box (64, 64)
    float* out;         // [M][128]
    int kiters;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) exp2cta_kernel(const __grid_constant__ Exp2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;              // kiters x 16 KB
    uint8_t* sB = smem + 64 * 1024;  // kiters x 8 KB
    __shared__ __align__(8) uint64_t full, done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&full, 2);   // one arrival per CTA of the pair (+ the bytes of both)
        ptx::mbar_init(&done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tslot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;

    if (threadIdx.x == 0) {
        // leader's barrier, addressed in the cluster window
        const uint32_t full_leader = mapa_u32(ptx::smem_u32(&full), 0);
        const uint32_t bytes = p.kiters * (16384 + 8192);
        if (rank == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(&full)), "r"(2 * bytes) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(full_leader) : "memory");
        }
        for (int k = 0; k < p.kiters; ++k) {
            asm volatile(
                "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(ptx::smem_u32(sA + k * 16384)), "l"(reinterpret_cast<uint64_t>(&p.a_map)), "r"(full_leader), "r"(k * 64),
                "r"(pair * 256 + (int)rank * 128), "r"(0)
                : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(ptx::smem_u32(sB + k * 8192)), "l"(reinterpret_cast<uint64_t>(&p.b_map)), "r"(full_leader), "r"(k * 64),
                "r"((int)rank * 64), "r"(0)
                : "memory");
        }
        if (rank == 0) {
            ptx::mbar_wait(&full, 0);
            ptx::tc_fence_after();
            const uint32_t idesc = ptx::make_idesc(1, 256, 128);
            for (int k = 0; k < p.kiters; ++k) {
                const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA + k * 16384));
                const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB + k * 8192));
                for (int j = 0; j < 4; ++j) {
                    const uint32_t accum = (k | j) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                        ::"r"(tmem), "l"(da + 2 * j), "l"(db + 2 * j), "r"(idesc), "r"(accum)
                        : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(ptx::smem_u32(&done)), "h"((uint16_t)3)
                         : "memory");
        }
    }
    __syncwarp();
    ptx::mbar_wait(&done, 0);
    ptx::tc_fence_after();
    const int row = pair * 256 + rank * 128 + warp * 32 + lane;
    for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.out[(long long)row * 128 + c + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

}  // namespace dxmi

extern "C" int dxmi_exp_2cta_gemm(const void* a, const void* b, float* out, int M, int K, void* stream) {
    using namespace dxmi;
    Exp2Params p;
    int r = make_mat_map(&p.a_map, a, K, M, 1, K, 0, 128);
    if (r) return r;
    r = make_mat_map(&p.b_map, b, K, 128, 1, K, 0, 64);
    if (r) return r;
    p.out = out;
    p.kiters = K / 64;
    cudaFuncSetAttribute(exp2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    exp2cta_kernel<<<2 * (M / 256), 128, 100 * 1024, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
