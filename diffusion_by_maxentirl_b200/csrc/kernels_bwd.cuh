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

// ---- small helpers of the U-Net backward (engine_train_unet.cu) -----------------------------------------------------------
// dst[b][c][r] = src[b][r][c]: src = B matrices of R rows x C columns, row stride ld_src, batch stride bs_src (elements); dst dense
void transpose_bf16_batched(const bf16* src, long long ld_src, long long bs_src, bf16* dst, int R, int C, int B, cudaStream_t st);
// softmax backward per row: dS = P * (dP - sum_j P_j dP_j) * scale;  P bf16 [rows,S], dP fp32 [rows,S] -> dS bf16 [rows,S]
void softmax_bwd_rows(const bf16* P, const float* dP, bf16* dS, long long rows, int S, float scale, cudaStream_t st);
// whole single-head attention backward for short sequences (S <= 64), one CTA per image: qkv bf16 [N,S,3C] (q | k | v),
// d_o bf16 [N,S,C] -> dqkv bf16 [N,S,3C] (dq | dk | dv);  P = softmax(scale q k^T) is recomputed
void attn_small_bwd(const bf16* qkv, const bf16* d_o, bf16* dqkv, int N, int S, int C, float scale, cudaStream_t st);
// out[n, 2i+1, 2j+1, :] = dy[n, i, j, :], zero elsewhere (data / weight gradient of the stride-2 pad-(0,1,0,1) Downsample conv
// become the standard pad-1 operators on this tensor)
void zero_insert2x(const bf16* dy, bf16* out, int N, int h, int w, int C, cudaStream_t st);
// nearest-2x upsample backward: out[n,i,j,:] = sum of the 2x2 block of dy [N,2h,2w,C]
void sumpool2(const bf16* dy, bf16* out, int N, int h, int w, int C, cudaStream_t st);
// dst[r, 0:C] (+)= src[r*ld_src + 0:C]  (gradient fan-in of skip connections / residuals); init = 1 overwrites
void accum_bf16(bf16* dst, const bf16* src, long long ld_src, long long rows, int C, int init, cudaStream_t st);
// out[n*ld_out + c] = sum_p x[n, p, c]   (x bf16 [N,HW,C]): gradient of the per-image temb row vector
void colsum_per_image(const bf16* x, int N, int HW, int C, float* out, int ld_out, cudaStream_t st);
// fp32 Linear backward: dW[o,k] = sum_n dy[n,o] * act(x[n,k]), db[o] = sum_n dy[n,o]   (act: 0 none, 2 exact silu)
void linear_bwd_w(const float* dy, int ld_dy, const float* x, int ld_x, int act_x, float* dW, float* db, int N, int O, int K, cudaStream_t st);
// dx[n,k] (+)= sum_o dy[n,o] * W[o,k]
void linear_bwd_x(const float* dy, int ld_dy, const float* W, float* dx, int ld_dx, int N, int O, int K, int accumulate, cudaStream_t st);
// Backward of a ROW-STACK of Linear layers that share one input x [N, K] (every ResBlock's temb_proj, unet_small.py:123): dy [N, ld_dy]
// holds the layers' output gradients side by side (layer l: columns off[l] .. off[l] + O[l]).  One launch each for all layers:
//   dW_l = dy_l^T x,  db_l = column sums of dy_l      (separate destination tensors)       and      dx = sum_l dy_l W_l
struct LinearStack {
    static constexpr int MAX = 48;  // (ImageNet-64 ADM: 36 ResBlocks)
    int n_layers;
    int off[MAX + 1];        // column offsets, off[n_layers] = total columns
    float* dW[MAX];          // [O_l, K] fp32 or null
    float* db[MAX];          // [O_l] or null
    const float* W[MAX];     // [O_l, K] fp32
};
void linear_stack_bwd_w(const float* dy, int ld_dy, const float* x, int ld_x, const LinearStack& ls, int N, int K, cudaStream_t st);
void linear_stack_bwd_x(const float* dy, int ld_dy, const LinearStack& ls, float* dx, int ld_dx, int N, int K, cudaStream_t st);
// d[i] *= silu'(x[i])
void silu_bwd_mul(float* d, const float* x, long long n, cudaStream_t st);
// conv_out (C -> 3, 3x3) backward helpers: w_t[c][o][tap] = w[o][c][8-tap] (fp32, feeds conv3x3_first as a 3 -> C convolution);
// grad[o][c][tap] = t[c][o][8-tap] (t = conv_first_wgrad of (g, d_eps));  gb[o] = sum over n, pixels of d_eps[n,o,:]
void conv_out_transpose_weights(const float* w, float* w_t, int C, cudaStream_t st);
void conv_out_wgrad_fix(const float* t, float* grad, int C, cudaStream_t st);
// ws: N * C floats
void sum_nchw_channels(const float* x, int N, int C, int HW, float* ws, float* out, cudaStream_t st);
// y = x * sigmoid(x) (exact expf)
void silu_f32(const float* x, float* y, long long n, cudaStream_t st);
// training-mode dropout (unet_small.py:126-127): x[i] *= keep(i) / (1 - p) with a counter-based keep mask that depends only on
// (seed, stream, element index) - the backward regenerates it instead of storing it.  mask_out (optional): the scaled mask itself.
void dropout_bf16(bf16* x, long long n, float p, unsigned long long seed, unsigned stream, bf16* mask_out, cudaStream_t st);

// ---- ADM / EDM U-Net backward helpers (engine_train_adm.cu; models/cm/unet.py under autograd) ------------------------------
// FiLM GroupNorm (use_scale_shift_norm, cm/unet.py:250-254): z = (xh gamma + beta)(1 + scale[n,c]) + shift[n,c].  group_norm_bwd
// (called with dgamma = dbeta = null) already produces dx with the per-image effective gamma and leaves (A, B)[n,c] in `ws`; this
// finishes the parameter side:  d_shift = A, d_scale = gamma B + beta A, dgamma = sum_n (1 + scale) B, dbeta = sum_n (1 + scale) A.
// film / d_film: [N, ld] fp32 with scale at column c and shift at column C + c.
void gn_bwd_film_params(const float* ws, int N, int HW, int C, const float* film, int film_ld, const float* gamma, const float* beta,
                        float* d_film, float* dgamma, float* dbeta, cudaStream_t st);
// P[r, :] = softmax(scores[r, :]) (fp32 [rows, S] -> bf16 [rows, S]); one warp per row, any S
void softmax_rows(const float* scores, bf16* P, long long rows, int S, cudaStream_t st);
// multi-head attention backward for short sequences (S <= 64, head dim d <= 64): qkv bf16 [N, S, 3C] with q | k | v blocks of C
// columns and head h at column h*d of its block; d_o bf16 [N, S, C] -> dqkv (same layout as qkv).  One CTA per (image, head).
void attn_small_bwd_heads(const bf16* qkv, const bf16* d_o, bf16* dqkv, int N, int heads, int S, int d, float scale, cudaStream_t st);
// label-embedding gradient (cm/unet.py:778-779, emb = emb + label_emb(y)): grad[y[n], :] += d_emb[n, :] in image order
// (deterministic: one thread per column walks the batch); grad [num_classes, D] is zeroed first
void embedding_grad(const float* d_emb, const long long* idx, float* grad, int N, int D, int num_classes, cudaStream_t st);
// x *= s (bf16, in place)
void scale_bf16(bf16* x, long long n, float s, cudaStream_t st);

}  // namespace dxmi
