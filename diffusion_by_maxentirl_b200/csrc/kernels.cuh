// HBM-bound and small CUDA-core kernels of the DxMI sampler path (everything that is not a dense contraction).
// Activations are NHWC bf16 unless stated; network inputs / outputs and sampler states are NCHW fp32 like the
// reference (models/DxMI/var_sampler.py, models/DxMI/openai_diffusion.py).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace dxmi {

typedef __nv_bfloat16 bf16;

// ---- weight packing -------------------------------------------------------------------------------------------
// OIHW (fp32 or fp16) conv weight -> bf16 [Cout, ldk] K-major rows, K index = k_off + tap * c_cnt + (c - c_off)
// for input channels c in [c_off, c_off + c_cnt).  tap = r * kw + s.
void pack_conv_weight(const void* w, int w_is_half, int Cout, int Cin, int kh, int kw, int c_off, int c_cnt, bf16* dst,
                      long long ldk, long long k_off, cudaStream_t st);
// data-gradient packing: dst[ci][k_off + tap' * Cout + co] = W[co][ci][taps-1-tap']  (transposed, taps flipped), so that
// dX = the forward implicit-GEMM convolution of dY with these rows (3x3 pad 1 stride 1, or 1x1)
void pack_conv_weight_dgrad(const void* w, int w_is_half, int Cout, int Cin, int taps, bf16* dst, long long ldk, long long k_off,
                            cudaStream_t st);
// OIHW 3x3 weight -> bf16 [4 phases][Cout][4 taps * Cin]: the pre-summed 2x2 filters of conv3x3(nearest_upsample2x(.)) (ConvGemmParams::up2)
void pack_conv_weight_up2(const void* w, int w_is_half, int Cout, int Cin, bf16* dst, cudaStream_t st);
void cast_to_f32(const void* src, int src_is_half, float* dst, long long n, cudaStream_t st);

// ---- first / last convolutions (3 <-> C channels; HBM bound) -------------------------------------------------------
// x fp32 NCHW [N,Cin<=4,H,W] (optionally scaled by in_scale[n]) -> bf16 NHWC [N,H,W,Cout]; w fp32 OIHW; act = GemmAct.
// stats (optional): GroupNorm partials of the output, [N][H*W/128][Cout][2] (sum, sumsq) like the GEMM epilogues write.
void conv3x3_first(const float* x, const float* in_scale, const float* w, const float* b, bf16* out, float* stats, int N,
                   int Cin, int H, int W, int Cout, int act, cudaStream_t st);
// the same operator on warp-level tensor-core MMAs (conv_first.cu; inference plans): x and w are rounded to bf16 like every other
// activation / weight of the path, accumulation in fp32
bool conv3x3_first_tc_supported(int Cin, int H, int W, int Cout);
void conv3x3_first_tc(const float* x, const float* in_scale, const float* w, const float* b, bf16* out, float* stats, int N, int H, int W,
                      int Cout, int act, cudaStream_t st);
// h bf16 NHWC [N,H,W,C] -> fp32 NCHW [N,Cout<=4,H,W]
void conv3x3_last(const bf16* h, const float* w, const float* b, float* out, int N, int C, int H, int W, int Cout,
                  cudaStream_t st);

// ---- GroupNorm(+SiLU)(+FiLM) ------------------------------------------------------------------------------------------
// Input = channel concat of up to two NHWC bf16 tensors (C1 + C2 channels, pixel strides ld1 / ld2).
// stats: partial[(n * slabs + s) * 64 + g * 2 + {0,1}] = (sum, sumsq) of group g over slab s.
// apply: y = (x - mean) * rstd * gamma + beta ; optional FiLM y = y * (1 + scale[n,c]) + shift[n,c] (film = [N, 2C]);
//        optional SiLU; output bf16 NHWC [N,HW,C1+C2].
int gn_num_slabs(int N, int HW);
void gn_stats(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups,
              float* partial, int slabs, cudaStream_t st);
void gn_apply(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups, float eps,
              const float* gamma, const float* beta, const float* film, int film_ld, int silu, const float* partial,
              int slabs, bf16* out, cudaStream_t st);

// apply with statistics from the producer GEMMs' fused partials st1 / st2 = [N][P1|P2][C1|C2][2] (gemm_tc2.cu)
void gn_apply_fused(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups, float eps,
                    const float* gamma, const float* beta, const float* film, int film_ld, int silu, const float* st1,
                    int P1, const float* st2, int P2, bf16* out, cudaStream_t st);
// two launches: per-image finalize (statistics -> per-channel affine, ab_ws = [N][C1+C2][2] floats) + streaming apply
void gn_finalize_apply(const bf16* x1, int C1, int ld1, const bf16* x2, int C2, int ld2, int N, int HW, int groups,
                       float eps, const float* gamma, const float* beta, const float* film, int film_ld, int silu,
                       const float* st1, int P1, const float* st2, int P2, float* ab_ws, bf16* out, cudaStream_t st,
                       float* mr = nullptr);  // mr (optional): [N][groups][2] (mean, rstd), kept by training plans
void set_gn_unroll(int v);  // 4 | 8 independent 16-byte loads per thread in the streaming apply kernel
// [N][P][C][2] -> [N][1][C][2]
void gn_collapse(const float* in, float* out, int N, int P, int C, cudaStream_t st);
// y = bf16(silu(x))  (A operand of the batched temb / emb projection GEMM)
void silu_to_bf16(const float* x, bf16* y, long long n, cudaStream_t st);

// ---- small dense helpers (fp32 SIMT) -----------------------------------------------------------------------------------
// sinusoidal features: out[n, :] = [sin | cos] (order 0, DDPM: freq = exp(-ln(1e4) * i / (half-1)))
//                                  [cos | sin] (order 1, ADM:  freq = exp(-ln(1e4) * i / half))
void timestep_embedding(const float* t, float* out, int N, int dim, int order, cudaStream_t st);
// y[n, o] = sum_k act_in(x[n, k]) * W[o, k] + b[o] (+ add[n, o]);  act_in: 0 none, 2 silu; act_out likewise.
void linear_f32(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int N, int K, int O,
                int act_in, int act_out, cudaStream_t st);
void embedding_add(float* emb, const float* table, const long long* idx, int N, int D, cudaStream_t st);

// ---- resampling ----------------------------------------------------------------------------------------------------------
void upsample2x(const bf16* x, bf16* out, int N, int H, int W, int C, cudaStream_t st);          // nearest
// last convolution of a U-Net (3x3, pad 1, C -> Cout <= 8) as a read-once stream: conv_last.cu.  wp: bf16 [8][9*C] (k = tap*C + c, rows >= Cout
// zero), bias8: fp32 [8]; out fp32 NCHW
bool conv3x3_last_supported(int H, int W, int C, int Cout);
struct ConvLastOp {
    CUtensorMap xmap;  // (c, w, h, n) over the input, box (64, W + 2, 128 / W + 2, 1): one halo tile of one 64-channel slice
    int N, H, W, C;
};
int prepare_conv3x3_last(const bf16* x, int N, int H, int W, int C, ConvLastOp* op);  // at plan-build time (the input pointer is fixed)
void conv3x3_last(const ConvLastOp& op, const bf16* wp, const float* bias8, float* out, int Cout, cudaStream_t st);
// out[n][c][p] = src[(n*HW + p)*ld + c] for c < C (fp32): the padded NHWC result of the network-output convolution -> NCHW
void nhwc_to_nchw_f32(const float* src, int ld, float* out, int N, int HW, int C, cudaStream_t st);
void avgpool2(const bf16* x, bf16* out, int N, int H, int W, int C, int act, cudaStream_t st);   // 2x2 mean (+act)

// ---- tiny attention (whole sequence in one CTA; seq <= 64) ------------------------------------------------------------
// q,k,v: bf16 [N, seq, ld] with head h at channel offset h*d (+ q_off/k_off/v_off); out bf16 [N, seq, ldo].
void attn_small(const bf16* q, const bf16* k, const bf16* v, int ld, bf16* out, int ldo, int N, int heads, int seq,
                int d, float scale, cudaStream_t st);

// ---- sampler transitions (fp32 NCHW) ------------------------------------------------------------------------------------
// VARSampler step (var_sampler.py:250-295 / :357-408): per-sample coefficients a,c,sigma [N].
//   mean = a*x + c*eps ; xn = mean + sigma*z ; control = c*eps ; logp[n] = mean_CHW N(xn; mean, sigma).log_prob
void var_step(const float* x, const float* eps, const float* z, const float* a, const float* c, const float* sigma,
              float* xn, float* mean, float* control, float* logp, uint8_t* u8, int N, int CHW, cudaStream_t st);
// EDM ancestral step (openai_diffusion.py:67-99, karras_diffusion.py:336-351): per-sample coefficient table
// coef[n] = {c_skip, c_out, sigma, sigma_down, sigma_noise}; F = raw network output.
//   D = c_out*F + c_skip*x ; mu = x + (x - D)/sigma * (sigma_down - sigma) ; xn = mu + sigma_noise * z
void edm_step(const float* x, const float* F, const float* z, const float* coef, float* xn, float* mean, uint8_t* u8, int N, int CHW,
              cudaStream_t st);

// broadcast one EDM schedule row {c_in, rescaled_t, c_skip, c_out, sigma, sigma_down} (host) + the step's noise scale
// (device scalar: it derives from the learnable log_betas) to per-sample tables
void edm_fill(float* coef, float* x_scale, float* t, int N, const float* row6, const float* sigma_noise_dev, cudaStream_t st);
// same for the VARSampler rollout: {tau, a, c} host scalars + device sigma
void var_fill(float* t, float* a, float* c, float* sigma, int N, float tau, float av, float cv, const float* sigma_dev,
              cudaStream_t st);

// ---- value head (modules.py:150-158): relu -> sum over HW -> Linear(C,1) -> Linear(1,1) ---------------------------------
void value_head(const bf16* h, int N, int HW, int C, const float* lin_w, const float* lin_b, const float* scale_w,
                const float* scale_b, float* out, cudaStream_t st);

// ---- tiny utilities ------------------------------------------------------------------------------------------------------
void fill_f32(float* p, float v, long long n, cudaStream_t st);
void vec_add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st);

// ---- layout helpers ----------------------------------------------------------------------------------------------------------
void nhwc_bf16_to_nchw_f32(const bf16* x, float* out, int N, int C, int HW, cudaStream_t st);
void nchw_f32_to_nhwc_bf16(const float* x, bf16* out, int N, int C, int HW, cudaStream_t st);
// samples in [-1,1] fp32 -> uint8 ((x+1)*127.5, clamp) (generate_large.py:43)
void quantize_u8(const float* x, uint8_t* out, long long n, cudaStream_t st);

}  // namespace dxmi
