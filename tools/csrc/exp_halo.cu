// EXPERIMENT (not on the product path): does a SWIZZLE_128B K-major UMMA descriptor whose start address is offset by a
// multiple of 128 bytes (not 1024-aligned) address the rows a TMA halo load wrote?  If yes, one halo tile
// [(bh+2) x (W+2) pixels x 64 channels] serves all 9 taps of a 3x3 convolution (A traffic / 9).
//   grid = 1 CTA. x: NHWC bf16 [1, H, W=32, 64]; w: packed bf16 [N=64][9*64] (tap-major); out fp32 [128][64]:
//   row p <-> padded-linear position (h = p / 34, w = p % 34) of output rows h0.., valid when w < 32.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace dxmi {

struct HaloParams {
    CUtensorMap a_map;  // (c, w, h, n) box (64, 34, 6, 1)
    CUtensorMap b_map;  // (k, rows) box (64, 64)
    float* out;
    int h0;
    int mode;  // 0: base_offset field = 0; 1: base_offset = (addr >> 7) & 7
};

__global__ void __launch_bounds__(128) halo_kernel(const __grid_constant__ HaloParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                 // 6 rows * 34 px * 128 B = 26112 B -> pad to 32 KB
    uint8_t* sB = smem + 32 * 1024;     // 9 taps * [64 rows x 128 B] = 72 KB
    __shared__ __align__(8) uint64_t full, done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&full, 1);
        ptx::mbar_init(&done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc(&tslot, 64);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(&full, 6 * 34 * 128 + 9 * 64 * 128);
        ptx::tma_load_4d(sA, &p.a_map, &full, 0, -1, p.h0 - 1, 0);
        for (int t = 0; t < 9; ++t) ptx::tma_load_3d(sB + t * 8192, &p.b_map, &full, t * 64, 0, 0);
        ptx::mbar_wait(&full, 0);
        ptx::tc_fence_after();
        constexpr uint32_t idesc = ptx::make_idesc(1, 128, 64);
        for (int t = 0; t < 9; ++t) {
            const int r = t / 3, s = t % 3;
            const uint32_t a_addr = ptx::smem_u32(sA) + (r * 34 + s) * 128;
            uint64_t da = ptx::make_kmajor_sw128_desc(a_addr);
            if (p.mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
            const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB + t * 8192));
            for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, (t | k) ? 1u : 0u);
        }
        ptx::umma_commit(&done);
    }
    __syncwarp();
    ptx::mbar_wait(&done, 0);
    ptx::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.out[row * 64 + c + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 64);
    }
}

}  // namespace dxmi

extern "C" int dxmi_exp_halo_conv(const void* x, int H, const void* w_packed, float* out, int h0, int mode, void* stream) {
    using namespace dxmi;
    HaloParams p;
    // activation map with a 34 x 6 box (stride 1): reuse make_act_map with bw = 34, bh = 6, bn = 1
    int r = make_act_map(&p.a_map, x, 64, 32, H, 1, 64, 64LL * 32, 64LL * 32 * H, 34, 6, 1, 1);
    if (r) return r;
    r = make_mat_map(&p.b_map, w_packed, 9 * 64, 64, 1, 9 * 64, 0, 64);
    if (r) return r;
    p.out = out;
    p.h0 = h0;
    p.mode = mode;
    cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    halo_kernel<<<1, 128, 112 * 1024, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
