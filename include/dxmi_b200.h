/*
 * dxmi_b200.h - C ABI of the B200-native DxMI sampler-rollout path (libdxmi_b200.so).
 *
 * The reference (swyoon/Diffusion-by-MaxEntIRL) has NO native / FFI boundary: its hot path is three Python
 * nn.Module-level interfaces.  This header is the boundary we introduce *underneath* those interfaces; every
 * entry point names the reference call it replaces (paths relative to the reference root).
 *
 * Conventions: plain C types only (no torch types); all pointers are CUDA device pointers unless marked "host";
 * every call returns 0 on success or a negative/cuda error code and never throws; dxmi_last_error() gives the
 * message; work is enqueued on the caller's stream and the library never calls cudaDeviceSynchronize();
 * a handle is bound to one device and is not thread-safe (one process per GPU, like the reference under torchrun).
 * Network inputs/outputs and sampler states are fp32 NCHW contiguous (the reference's layout); weights are
 * borrowed by pointer in the reference's state_dict layout (OIHW fp32/fp16 convs, [out,in] linears) and re-packed
 * to bf16 K-major on the device by dxmi_finalize()/dxmi_repack().
 */
#ifndef DXMI_B200_H
#define DXMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dxmi_net_s* dxmi_net_t;
typedef void* dxmi_stream_t; /* cudaStream_t */

enum dxmi_arch_kind {
    DXMI_ARCH_DDPM_UNET = 0, /* models/DxMI/unet_small.py:194-332  Model            */
    DXMI_ARCH_ADM_UNET = 1,  /* models/cm/unet.py:523-790          UNetModel        */
    DXMI_ARCH_IGEBM_V2 = 2   /* models/modules.py:104-163          IGEBMEncoderV2   */
};

enum dxmi_dtype { DXMI_F32 = 0, DXMI_F16 = 1, DXMI_I64 = 2 };

/* Constructor arguments of the three reference networks (configs/.../T*.yaml), flattened. */
typedef struct {
    int arch;               /* dxmi_arch_kind */
    int resolution;         /* DDPM: resolution; ADM: image_size; IGEBM: input H=W */
    int in_channels;
    int out_channels;
    int ch;                 /* DDPM: ch; ADM: model_channels; IGEBM: nh */
    int n_levels;
    int ch_mult[8];
    int num_res_blocks;
    int n_attn;             /* number of entries in attn_resolutions */
    int attn_resolutions[8];/* DDPM: feature-map sizes (16); ADM: downsample rates ds (image_size // res) */
    int num_classes;        /* ADM: 0 = unconditional */
    int num_head_channels;  /* ADM: 64 */
    int num_heads;          /* ADM: used when num_head_channels == -1 */
    int use_scale_shift_norm;
    int resblock_updown;
    int learn_out_scale;    /* IGEBM: out_scale Linear(1,1) present */
    int precision;          /* 0: bf16 tcgen05 path (rel-L2 <= 2e-2 vs the reference); 1: fp32 mode - CUDA-core FFMA kernels on fp32
                               activations, rel-L2 <= 1e-5 per step (DDPM U-Net and IGEBM, the reference's fp32 networks) */
} dxmi_arch_desc;

/* -------------------------------------------------------------------------------------------- lifecycle */
/* Replaces nn.Module construction (generate_cifar10.py:143, script_util.py:104-158, value: T10.yaml:20-31). */
int dxmi_create(const dxmi_arch_desc* desc, int device, dxmi_net_t* out);
void dxmi_destroy(dxmi_net_t net);
/* Replaces load_state_dict()/parameters(): borrow one state_dict tensor (generate_cifar10.py:148-149). */
int dxmi_bind_weight(dxmi_net_t net, const char* key, const void* dev_ptr, int dtype, const int64_t* shape, int ndim);
/* Number of state_dict keys the architecture expects; key i is written to buf (for the host-side mirror). */
int dxmi_num_weights(dxmi_net_t net);
int dxmi_weight_key(dxmi_net_t net, int i, char* buf, int buflen);
int dxmi_weight_shape(dxmi_net_t net, int i, int64_t* shape, int* ndim);
/* Verify every expected key is bound and pack weights (bf16, K-major). Call again (or dxmi_repack) after an
 * optimizer step changed the borrowed tensors (trainer.py:389). */
int dxmi_finalize(dxmi_net_t net, dxmi_stream_t stream);
int dxmi_repack(dxmi_net_t net, dxmi_stream_t stream);

/* -------------------------------------------------------------------------------------------- networks */
/* unet_small.Model.forward(x, t) (unet_small.py:292-332) and UNetModel.forward(x, timesteps, y) (cm/unet.py:761-790).
 * x [B,Cin,H,W] fp32, t [B] fp32, y [B] int64 or NULL, out [B,Cout,H,W] fp32.  x_scale: optional [B] fp32 factor
 * applied to x on load (EDM c_in, karras_diffusion.py:349) or NULL. */
int dxmi_unet_forward(dxmi_net_t net, const float* x, const float* x_scale, const float* t, const int64_t* y,
                      float* out, int B, dxmi_stream_t stream);
/* TimeIndependentValue.forward(x, t) -> IGEBMEncoderV2.forward (value.py:8-12, modules.py:142-163). out [B] fp32. */
int dxmi_value_forward(dxmi_net_t net, const float* x, float* out, int B, dxmi_stream_t stream);

/* -------------------------------------------------------------------------------------------- transitions */
/* One VARSampler transition given eps (var_sampler.py:262-289 / :375-404). a, c, sigma: per-sample [B] fp32.
 * mean / control / logp may be NULL. */
int dxmi_var_step(const float* x, const float* eps, const float* z, const float* a, const float* c,
                  const float* sigma, float* x_next, float* mean, float* control, float* logp, int B, int chw,
                  dxmi_stream_t stream);
/* One OpenAIDiffusion transition given the raw net output F (openai_diffusion.py:75-94, karras_diffusion.py:350).
 * coef [B,5] fp32 = {c_skip, c_out, sigma, sigma_down, sigma_noise}. mean may be NULL. */
int dxmi_edm_step(const float* x, const float* F, const float* z, const float* coef, float* x_next, float* mean,
                  int B, int chw, dxmi_stream_t stream);

/* -------------------------------------------------------------------------------------------- rollouts */
/* VARSampler.sample() (var_sampler.py:411-428 -> VAR_sampling :204-297): the whole T-step loop in one call.
 * sched (host) [T,3] fp32 rows {tau_i, a_i, c_i * adhoc_scale1} (Appendix E.1 of SURVEY.md); sigma (DEVICE) [T] fp32 =
 * the per-step noise scale (it derives from the learnable log_betas, so it stays on the device: no host sync);
 * noise [T+1,B,C,H,W] fp32: noise[0] = x_0, noise[1+i] = z of step i (host-supplied noise is part of the parity
 * contract); l_sample [T+1,B,C,H,W]; mean, control [T,B,C,H,W] or NULL; logp [T,B] or NULL.
 * sample_u8 [B,C,H,W] or NULL: the final samples quantised to uint8 by the LAST transition kernel itself
 * (((x+1)*127.5).clamp(0,255), generate_cifar10.py:205-209 / generate_large.py:43) - the post-rollout step fused in. */
int dxmi_var_rollout(dxmi_net_t net, const float* sched_host, const float* sigma_dev, int T, const float* noise,
                     float* l_sample, float* mean, float* control, float* logp, uint8_t* sample_u8, int B,
                     dxmi_stream_t stream);
/* OpenAIDiffusion.sample() (openai_diffusion.py:101-127). sched (host) [T,6] fp32 rows
 * {c_in, rescaled_t, c_skip, c_out, sigma, sigma_down}; sigma_noise (DEVICE) [T] fp32 = the noise scale actually applied
 * per step (openai_diffusion.py:79-92); noise[0] = x_0 (already scaled by sigma_max). sample_u8: as above (or NULL). */
int dxmi_edm_rollout(dxmi_net_t net, const float* sched_host, const float* sigma_noise_dev, int T, const float* noise,
                     const int64_t* y, float* l_sample, float* mean, uint8_t* sample_u8, int B, dxmi_stream_t stream);

/* -------------------------------------------------------------------------------------------- trainer-side fused ops (8f rank 1) */
/* DxMI_Trainer.get_running_cost (trainer.py:163-169): rc[n] = mean_CHW((next_state - state)^2) / (2 beta_next[n]); all fp32,
 * beta_next [B] = betas_for_q gathered at n_timesteps - t - 1 by the caller. */
int dxmi_running_cost_fwd(const float* state, const float* next_state, const float* beta_next, float* rc, int B, int chw,
                          dxmi_stream_t stream);
/* its backward: d_next_state = grad_rc[n] (next_state - state) / (beta_next[n] CHW), d_state = -that (either may be NULL) */
int dxmi_running_cost_bwd(const float* state, const float* next_state, const float* beta_next, const float* grad_rc,
                          float* d_next_state, float* d_state, int B, int chw, dxmi_stream_t stream);
/* Multi-tensor optimizer step (torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step(), trainer.py:324-327 / :388-389,
 * train_cifar10.py:283-296). One device table row per fp32 parameter tensor; the caller splits every tensor into chunks of
 * dxmi_opt_chunk_elems() elements: chunk_tensor[k] = table row, chunk_first[k] = first element of chunk k in that tensor. */
typedef struct dxmi_opt_tensor {
    float* param;
    float* grad;       /* NULL: tensor skipped (no gradient this step) */
    float* exp_avg;
    float* exp_avg_sq;
    long long numel;
    float lr;          /* per tensor: the two-LR parameter groups of train_cifar10.py:287-290 */
    int reserved;
} dxmi_opt_tensor;
int dxmi_opt_chunk_elems(void);
/* global L2 norm of all gradients -> norm_coef_out[0] = norm, [1] = min(1, max_norm / (norm + 1e-6)) (device, no host sync);
 * partial_ws: n_chunks floats. scale_grads != 0 also multiplies the gradients in place like clip_grad_norm_ does. */
int dxmi_opt_grad_norm(const dxmi_opt_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_first_dev, int n_chunks,
                       float max_norm, float* partial_ws, float* norm_coef_out, int scale_grads, dxmi_stream_t stream);
/* Adam (weight_decay 0, amsgrad off) on grad * coef, coef = norm_coef[1] (or 1 when NULL); step counts from 1 */
int dxmi_opt_adam_step(const dxmi_opt_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_first_dev, int n_chunks,
                       const float* norm_coef_or_null, float beta1, float beta2, float eps, int step, int scale_grads_in_place,
                       dxmi_stream_t stream);

/* samples in [-1,1] -> uint8 ((x+1)*127.5 clamp), generate_large.py:43 */
int dxmi_quantize_u8(const float* x, uint8_t* out, long long n, dxmi_stream_t stream);

/* -------------------------------------------------------------------------------------------- kernel-level ops
 * (exposed for parity tests and profiling of the individual kernels) */
typedef struct {
    /* A: up to three NHWC bf16 sources sharing the same (N, H, W) */
    const void* a_ptr[3];
    int a_C[3];              /* channels used from each source */
    int a_ld[3];             /* pixel stride in elements */
    int N, H, W;             /* input images / spatial size */
    int nseg;
    int seg_src[3], seg_taps[3]; /* source index and 1|9 taps per K segment */
    int stride;              /* 1 | 2 (2 = Downsample with pad (0,1,0,1), unet_small.py:69-73) */
    int out_H, out_W;
    /* B: packed bf16 [rows, K_total] (+ batch) */
    const void* b_ptr;
    int b_rows;
    long long b_ld;
    long long b_batch_stride;
    int batch;               /* grid.z; 1 for convs */
    int a_batched, b_batched;
    /* epilogue */
    void* out;
    int ldo;
    long long out_batch_stride;
    int out_fp32;
    const float* bias;
    int bias_along_m;
    const float* rowvec;
    int ldrv;
    int rows_per_image;
    const void* residual;
    int ldr;
    long long res_batch_stride;
    int act;                 /* 0 none, 1 leaky-relu(0.2), 2 silu */
    float alpha;
    int softmax;
    int block_n;             /* 0 = auto */
    int out_nchw;            /* fp32 output in [image][b_rows][rows_per_image] order (network output layout) */
    float* gn_stats;         /* optional: GroupNorm partials of the bf16 output, [rows/gn_seg][b_rows][2] (sum, sumsq) */
    int gn_seg;              /* rows per partial: 32, 64 or 128; must divide the rows of one image */
    int gn_halo_P;           /* > 0: the partials buffer is [N][gn_halo_P][b_rows][2], one partial per halo tile (3x3 convs on
                                32/64-wide maps, see dxmi_op_halo_tiles_per_image); 0: row-segment partials */
    const void* gate;        /* optional bf16 [rows, ldg]: out *= (gate > 0 ? 1 : 0.2) after everything else - the backward of
                                leaky-relu(0.2) through the saved activation (modules.py:81,99 under autograd) */
    int ldg;
    int up2;                 /* 1: nearest-2x upsample folded into the 3x3 convolution that follows it (unet_small.py:52-64 Upsample,
                                cm/unet.py:103-118, ResBlock up=True :186-199) as four 2x2 phase convolutions of the LOW-resolution
                                input: N/H/W/out_H/out_W = the low-resolution geometry, one 4-tap segment, batch = 4 (phase = py*2+px),
                                b_ptr = dxmi_op_pack_conv_weight_up2 rows [4][b_rows][4*C] with b_batched = 1; `out` is the
                                [N, 2H, 2W, ldo] tensor; gn_stats partials are [N][4][H*W/gn_seg][b_rows][2] */
} dxmi_gemm_desc;

int dxmi_op_conv_gemm(const dxmi_gemm_desc* d, dxmi_stream_t stream);
int dxmi_op_pack_conv_weight(const void* w, int dtype, int Cout, int Cin, int kh, int kw, int c_off, int c_cnt,
                             void* dst_bf16, long long ldk, long long k_off, dxmi_stream_t stream);
/* OIHW 3x3 weight -> the four pre-summed 2x2 phase filters of conv3x3(nearest_upsample2x(x)): bf16 [4][Cout][4*Cin], phase =
 * py*2+px, k = (dy*2+dx)*Cin + c; rows ky in {0},{1,2} (py = 0) or {0,1},{2} (py = 1) are summed in fp32, likewise columns. */
int dxmi_op_pack_conv_weight_up2(const void* w, int dtype, int Cout, int Cin, void* dst_bf16, dxmi_stream_t stream);
int dxmi_op_group_norm(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int N, int HW, int groups,
                       float eps, const float* gamma, const float* beta, const float* film, int film_ld, int silu,
                       float* partial_ws, void* out, dxmi_stream_t stream);
int dxmi_op_gn_ws_floats(int N, int HW, int groups);
/* tiles per image of the halo-mode 3x3 convolution on an H x W map (0: that geometry uses plain 128-pixel tiles) */
int dxmi_op_halo_tiles_per_image(int H, int W);
/* Fused d=64 multi-head attention (QKVAttentionLegacy.forward, models/cm/unet.py:413-441): qk bf16 [B, seq, ld_qk] with
 * head h's queries at column q_col0 + 64h and keys at k_col0 + 64h; vt bf16 [B, heads*64, seq] (V transposed);
 * out bf16 [B, seq, ldo], head h at column 64h; scale multiplies the logits (d^-1/2). seq % 128 == 0. */
int dxmi_op_attention(const void* qk, long long ld_qk, int q_col0, int k_col0, const void* vt, void* out, int ldo, int B,
                      int heads, int seq, float scale, dxmi_stream_t stream);

/* ---- value-net training (trainer.py:244-264 energy update, :276-326 TD updates, :369-389 value term of the
 * sampler loss): what `loss.backward()` does underneath TimeIndependentValue.  bf16 mode only. ---- */
/* fp32 gradient buffer (state_dict shape) for one key; the backward WRITES it (the caller accumulates); NULL = skip this key. */
int dxmi_bind_grad(dxmi_net_t net, const char* key, float* dev_ptr);
/* dxmi_value_forward that keeps the activations of this batch for one dxmi_value_backward call (same B, same x). */
int dxmi_value_forward_train(dxmi_net_t net, const float* x, float* out, int B, dxmi_stream_t stream);
/* dout [B] fp32 = d loss / d out; fills every bound gradient and, if dx != NULL, dx [B,3,H,W] fp32 = d loss / d x. */
int dxmi_value_backward(dxmi_net_t net, const float* x, const float* dout, float* dx, int B, dxmi_stream_t stream);

/* ---- DDPM U-Net training (trainer.py:348-389 update_sampler: eps = net(x, t) under autograd), including training-mode dropout (unet_small.py:126-127; counter-based masks).
 * dxmi_unet_forward_train = dxmi_unet_forward (DDPM, no x_scale / labels) that keeps the activations of this batch; dxmi_unet_backward
 * takes dout [B,3,H,W] fp32 = d loss / d eps and writes every gradient bound with dxmi_bind_grad (no input gradient: the sampler
 * update detaches the state). ---- */
int dxmi_unet_forward_train(dxmi_net_t net, const float* x, const float* t, float* out, float dropout_p,
                            unsigned long long dropout_seed, int B, dxmi_stream_t stream);
/* The scaled keep mask (0 or 1/(1-p), bf16, NHWC order of swish(norm2(h))) that ResnetBlock number `stream_id` (position on the
 * forward tape: conv_in = 0, then blocks / attention / resampling in execution order) applies for (p, seed) - lets a reference
 * implementation replay the exact dropout pattern (tests). */
int dxmi_op_dropout_mask(void* mask_bf16, long long n, float p, unsigned long long seed, unsigned stream_id, dxmi_stream_t stream);
/* ---- ADM / EDM U-Net training (trainer.py:693-746 update_sampler_mixed_precision: F = net(c_in x, c_noise, y)
 * under autograd, models/cm/unet.py:761-790).  dxmi_adm_forward_train = dxmi_unet_forward for the ADM U-Net (labels y or NULL)
 * that keeps the activations of this batch; the gradients come from dxmi_unet_backward (dx must be NULL: the sampler update
 * differentiates one step from a replay-buffer state, which is a leaf). ---- */
int dxmi_adm_forward_train(dxmi_net_t net, const float* x, const float* t, const int64_t* y, float* out, float dropout_p,
                           unsigned long long dropout_seed, int B, dxmi_stream_t stream);
int dxmi_unet_backward(dxmi_net_t net, const float* x, const float* dout, float* dx /* [B,Cin,H,W] fp32 gradient w.r.t. the input state, or NULL */,
                       int B, dxmi_stream_t stream);

/* ---- backward operators (the training step differentiates through the value net, trainer.py:252-264,
 * :320-326, :369-389) ---- */
/* Packs an OIHW conv weight for the DATA gradient: dst[ci][k_off + tap' * Cout + co] = W[co][ci][taps-1-tap'] (bf16), so that
 * dX = dxmi_op_conv_gemm(dY, these rows) for a 3x3 pad-1 stride-1 (taps = 9) or 1x1 (taps = 1) convolution. */
int dxmi_op_pack_conv_weight_dgrad(const void* w, int dtype, int Cout, int Cin, int taps, void* dst_bf16, long long ldk,
                                   long long k_off, dxmi_stream_t stream);
/* WEIGHT gradient on the tensor cores: grad[co][ci_off + ci][tap] = scale * sum_p dY[p, co] * X[p + tap, ci]  (fp32 OIHW
 * tensor with Cin_total input channels; dy bf16 NHWC [N,H,W,Cout], x bf16 NHWC [N,H,W,Cin]; Cout % 128 == 0,
 * Cin in {64,128,192,256}; taps 9 (3x3, pad 1) or 1).  ws: dxmi_op_wgrad_ws_floats(...) fp32 scratch (split-K partials,
 * reduced in a fixed order: deterministic). */
long long dxmi_op_wgrad_ws_floats(int N, int H, int W, int Cout, int Cin, int taps);
int dxmi_op_conv_wgrad(const void* dy, const void* x, int N, int H, int W, int Cout, int Cin, int taps, float* grad_oihw,
                       int Cin_total, int ci_off, float scale, float* ws, dxmi_stream_t stream);

/* GroupNorm(32)(+SiLU) backward (groundwork for the U-Net backward; unet_small.py:119-126 under autograd).  x = concat(x1 | x2)
 * NHWC bf16, dy [N,HW,C] bf16, ab [N][C][2] = the forward's per-(image, channel) affine (a = rstd*gamma, b = beta - mean*a), mr
 * [N][32][2] = (mean, rstd).  Writes dx [N,HW,C] bf16 and, when not NULL, dgamma / dbeta [C] fp32.  ws: dxmi_op_gn_bwd_ws_floats. */
long long dxmi_op_gn_bwd_ws_floats(int N, int HW, int C);
int dxmi_op_group_norm_bwd(const void* x1, int C1, const void* x2, int C2, const void* dy, const float* ab, const float* mr, int N, int HW,
                           int groups, int silu, float* ws, void* dx, float* dgamma, float* dbeta, dxmi_stream_t stream);

/* -------------------------------------------------------------------------------------------- misc */
const char* dxmi_last_error(void);
/* A/B switches read when a plan is built (unless noted): "up2" (1, default: nearest-2x upsample + 3x3 conv as four 2x2 phase
 * convolutions of the low-resolution tensor; 0 = upsample2x_k + 9-tap convolution), "first_tc" (1, default: first convolution of the inference plans on
 * mma.sync tiles; 0 = fp32 FMA kernel), "lean_epi" (0, default; 1 = unstaged drain of bias-only GEMMs, measured slower), "attnblk" (1 = fused DDPM AttnBlock kernel), "pair"
 * / "pair_min" (cta_group::2 GEMM: 0 off, 1 when >= pair_min pair tiles, 2 always), "shift3", "s3_m2", "wave_bn", "small_map_bn",
 * "conv_out_padded", "gn_fused" (0 = one-kernel GroupNorm for maps up to 8x8, 1 = every map, N > 1 = maps up to N pixels), "gn_unroll",
 * "stats16", "s3_stages_max", "pair_resident_b" (weights-stationary pair GEMM, measured slower), "t_uniform" (1, default: one-row timestep
 * embedding in rollouts), "rollout_split" / "rollout_split_min" (batch-split rollouts on side streams, measured neutral), "pdl" (read at
 * every launch), and the ones listed at dxmi_set_timing_dump below */
int dxmi_set_option(const char* name, int value);
/* profiling only: device buffer of 8 int64 per CTA that the GEMM kernel fills with per-phase globaltimer stamps (NULL = off) */
int dxmi_set_debug_buffer(void* dev_ptr);
/* profiling only: with "time_gemms" on, dxmi_gemm_timing() also appends one CSV row per GEMM launch to this file (NULL = off) */
int dxmi_set_timing_dump(const char* path); /* "block_n_256" (tile width), "time_gemms" (event-time every GEMM), "gemm_version" (1 = one tile per CTA,
 * 2 = persistent kernel, default), "halo" (1 = halo-tile A reuse for 3x3 convs; default 0, measured slower), "dbg_mode" (profiling) */
/* with "time_gemms" on: summed CUDA-event duration / algorithmic FLOPs / count of the tcgen05 GEMM launches since
 * the last call (synchronises on the recorded events) */
int dxmi_gemm_timing(double* ms_total, double* flops_total, long long* launches);
/* with "time_gemms" on: the same for the HBM-bound kernel families, with their ALGORITHMIC bytes (one read + one write of the
 * tensor in its storage type). category 1 = GroupNorm finalize + apply, 2 = transition step x' = mu + sigma z (K6) */
int dxmi_aux_timing(int category, double* ms_total, double* bytes_total, long long* launches);
/* algorithmic 2*M*N*K FLOPs of all tcgen05 GEMM launches of one forward at batch B (0 if that plan is not built) */
double dxmi_plan_gemm_flops(dxmi_net_t net, int B);
/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
long long dxmi_launch_count(void);
size_t dxmi_workspace_bytes(dxmi_net_t net, int B);

#ifdef __cplusplus
}
#endif
#endif /* DXMI_B200_H */
