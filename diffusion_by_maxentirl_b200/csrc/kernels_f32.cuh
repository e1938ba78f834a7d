// fp32 mode (BASELINE.json north_star: "relative L2 at most 1e-5 per step in fp32 mode"): CUDA-core FFMA kernels with fp32
// NHWC activations and the borrowed fp32 state_dict tensors used as they are (OIHW, no packing).  This mode exists to
// validate the path against the reference's fp32 arithmetic (DDPM U-Net + value net); the
// throughput path is the bf16 tcgen05 one.  Every reduction runs in a fixed order (deterministic, batch invariant).
#pragma once
#include <cuda_runtime.h>

namespace dxmi {

struct ConvF32 {
    const float* x1 = nullptr;  // NHWC [N,H,W,C1]
    const float* x2 = nullptr;  // optional channel concat [N,H,W,C2]  (torch.cat([h, skip], dim=1))
    int C1 = 0, C2 = 0;
    const float* w = nullptr;   // OIHW [Cout, C1+C2, k, k]
    const float* bias = nullptr;
    const float* rowvec = nullptr;  // optional [N, ldrv]: per-image vector added to every pixel (temb projection)
    int ldrv = 0;
    const float* residual = nullptr;  // optional NHWC [N,Ho,Wo,Cout]
    float* out = nullptr;       // NHWC [N,Ho,Wo,Cout]
    int N = 0, H = 0, W = 0, Cout = 0;
    int k = 3;                  // 1 or 3
    int stride = 1;             // 1: zero padding (k-1)/2;  2: pad right / bottom by one, no other padding (Downsample)
    int act = 0;                // 0 none, 1 leaky-relu 0.2 (after bias / residual)
};
void conv_f32(const ConvF32& c, cudaStream_t st);

// GroupNorm over the channel concat of x1 | x2 (NHWC fp32), centred two-pass variance, optional x*sigmoid(x).
void group_norm_f32(const float* x1, int C1, const float* x2, int C2, int N, int HW, int groups, float eps,
                    const float* gamma, const float* beta, int silu, float* out, cudaStream_t st);
// single-head attention on NHWC fp32 q, k, v [N, HW, C]:  out = softmax(q k^T * scale) v
void attention_f32(const float* q, const float* k, const float* v, float* out, int N, int HW, int C, float scale,
                   cudaStream_t st);
// y[n, o] = sum_k act_in(x[n, k]) W[o, k] + b[o];  act_in 0 none, 2 x*sigmoid(x) (exact expf)
void linear_exact_f32(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int N, int K, int O,
                      int act_in, cudaStream_t st);
void upsample2x_f32(const float* x, float* out, int N, int H, int W, int C, cudaStream_t st);
void avgpool2_f32(const float* x, float* out, int N, int H, int W, int C, int act, cudaStream_t st);
void nchw_to_nhwc_f32(const float* x, float* out, int N, int C, int HW, cudaStream_t st);
void nhwc_to_nchw_f32(const float* x, float* out, int N, int C, int HW, cudaStream_t st);
// modules.py:150-158: relu -> sum over HW -> Linear(C,1) -> Linear(1,1)
void value_head_f32(const float* h, int N, int HW, int C, const float* lin_w, const float* lin_b, const float* scale_w,
                    const float* scale_b, float* out, cudaStream_t st);

}  // namespace dxmi
