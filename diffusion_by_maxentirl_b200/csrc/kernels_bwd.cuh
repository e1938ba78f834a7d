// Elementwise / reduction kernels of the value-net backward pass (engine_train.cu).  All reductions run in
// a fixed order (deterministic).
#pragma once
#include "kernels.cuh"

namespace dxmi {

// modules.py:150-158 backward.  h bf16 [N,HW,C] (the last block's activated output), dout fp32 [N].
//   g[n] = dout[n] * (scale_w ? scale_w[0] : 1);  dz[n,p,c] = g[n] * lin_w[c] * (h > 0)   (relu o leaky-relu: slope 1 where h > 0, else 0)
//   S[n,c] = sum_p relu(h[n,p,c])  (fp32 scratch for the parameter gradients)
void value_head_bwd(const bf16* h, const float* dout, const float* lin_w, const float* scale_w, bf16* dz, float* S, int N, int HW,
                    int C, cudaStream_t st);
// d lin_w[c] = sum_n g[n] S[n,c];  d lin_b = sum_n g[n];  d scale_w = sum_n dout[n] * pre[n];  d scale_b = sum_n dout[n]
// with pre[n] = sum_c S[n,c] lin_w[c] + lin_b.  Any gradient pointer may be null (not bound).
void value_head_param_grads(const float* S, const float* dout, const float* lin_w, const float* lin_b, const float* scale_w,
                            float* g_lin_w, float* g_lin_b, float* g_scale_w, float* g_scale_b, int N, int C, cudaStream_t st);
// 2x2 average-pool backward: dx[n, y, x, c] = 0.25 * dy[n, y/2, x/2, c]   (dy is already gated by the leaky-relu after the pool)
void avgpool2_bwd(const bf16* dy, bf16* dx, int N, int H, int W, int C, cudaStream_t st);
// out[c] = sum over rows of x[row, c] (bf16 [rows, C] -> fp32 [C]); ws: colsum_ws_floats(rows, C) floats
long long colsum_ws_floats(long long rows, int C);
void colsum_bf16(const bf16* x, long long rows, int C, float* ws, float* out, cudaStream_t st);
// first convolution (Cin = 3, fp32 NCHW input x) weight gradient: grad[co][ci][tap] = sum_p dz[p, co] * x[p + tap, ci]
// ws: N * (H/4) * Cout * 27 floats (one partial per 4-row slab of an image, summed in a fixed order); H % 4 == 0
void conv_first_wgrad(const bf16* dz, const float* x, float* ws, float* grad, int N, int H, int W, int Cout, cudaStream_t st);

// ---- GroupNorm(+SiLU) backward (groundwork for the U-Net backward) -------------------------
// Forward:  xh = (x - mean[n,g]) * rstd[n,g];  z = a[n,c] * x + b[n,c]  (a = rstd * gamma, b = beta - mean * a);  y = silu ? z*sigmoid(z) : z.
// Inputs: x = channel concat of x1 | x2 (NHWC bf16), dy [N,HW,C] bf16, ab [N][C] float2 (the forward's affine), mr [N][groups] float2
// (mean, rstd).  Outputs: dx [N,HW,C] bf16 (concat layout), dgamma / dbeta [C] fp32 (written; null = skip).
//   dz = dy * silu'(z);  A[n,c] = sum_p dz;  B[n,c] = sum_p dz * xh;  per group: SA = sum_c gamma_c A, SB = sum_c gamma_c B, m = cpg*HW
//   dx = rstd * (gamma_c dz - SA/m - xh SB/m);   dgamma_c = sum_n B[n,c];  dbeta_c = sum_n A[n,c]
// ws: gn_bwd_ws_floats(N, HW, C) floats.  All reductions in a fixed order.
long long gn_bwd_ws_floats(int N, int HW, int C);
void group_norm_bwd(const bf16* x1, int C1, const bf16* x2, int C2, const bf16* dy, const float* ab, const float* mr, int N, int HW,
                    int groups, int silu, float* ws, bf16* dx, float* dgamma, float* dbeta, cudaStream_t st);

}  // namespace dxmi
