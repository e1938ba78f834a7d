// Fused d=64 multi-head attention on tcgen05 (see attn_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gemm_tc.cuh"

namespace dxmi {

struct AttnParams {
    CUtensorMap qk_map;  // (cols, rows = seq, batch) bf16, box 64 x 128: q at column q_col0 + head*64, k at k_col0 + head*64
    CUtensorMap vt_map;  // (keys = seq, rows = C, batch) bf16, box 64 x 64: V^T rows head*64 ..
    __nv_bfloat16* out;  // [batch, seq, ldo], head h at column h*64
    int ldo;
    int seq;
    int q_col0, k_col0;
    int v_col0;          // >= 0: V is read from the SAME [batch, seq, ld] tensor at column v_col0 + head*64 (box 64 x 128, keys as
                         // rows) and enters P.V as an MN-major B operand - no separate V^T GEMM; < 0: vt_map holds V^T
    float scale_log2;    // softmax scale * log2(e)
};

struct AttnOp {
    AttnParams p;
    dim3 grid;
    double flops;
};

// Fused single-head d = 256, seq = 256 attention (attn256_tc.cu): the DDPM AttnBlock at 16x16.
struct Attn256Params {
    CUtensorMap qk_map_q;  // [B, 256, 512] bf16, box 64 x 128 (queries at columns 0..255)
    CUtensorMap qk_map_k;  // same tensor, box 64 x 256 (keys at columns 256..511)
    CUtensorMap vt_map;    // [B, 256 channels, 256 keys] bf16, box 64 x 256
    __nv_bfloat16* out;    // [B, 256, ldo]
    int ldo;
    float scale_log2;
};
struct Attn256Op {
    Attn256Params p;
    dim3 grid;
    double flops;
};
int prepare_attn256(const void* qk, const void* vt, void* out, int ldo, int B, float scale, Attn256Op* op);
int run_attn256(const Attn256Op& op, cudaStream_t st);

// Whole DDPM AttnBlock (GroupNorm -> q,k,v -> softmax(q k^T) v -> proj_out + residual) at seq 256, C = 256 (attnblk_tc.cu)
struct AttnBlkParams {
    CUtensorMap x_map;       // [B, 256 pixels, 256 channels] bf16, box 64 x 128
    CUtensorMap w_map;       // k | v | q | proj_out weights stacked as [1024, 256] bf16 (K-major), box 64 x 128
    const __nv_bfloat16* x;  // block input (residual)
    __nv_bfloat16* out;      // [B, 256, 256]
    const float* stats_in;   // GroupNorm partials of x from its producer: [B][P_in][256][2] (sum, sumsq)
    int P_in;
    float* stats_out;        // GroupNorm partials of out: [B][2][256][2] (one per 128-row half), or null
    const float* gamma;      // norm.weight / norm.bias
    const float* beta;
    const float* bias;       // k | v | q | proj_out biases, [1024]
    float eps;
    float scale_log2;
    long long* dbg;          // optional phase stamps [CTA][16] (globaltimer ns), tools/attnblk_timeline.py
};
void set_attnblk_dbg(void* p);
long long* attnblk_dbg_buffer();
struct AttnBlkOp {
    AttnBlkParams p;
    int B;
    double flops;
};
int prepare_attnblk256(const void* x, const void* w_kvqp, const float* bias_kvqp, const float* gamma, const float* beta,
                       const float* stats_in, int P_in, float eps, float scale, void* out, float* stats_out, int B, AttnBlkOp* op);
int run_attnblk256(const AttnBlkOp& op, cudaStream_t st);

// vt == nullptr: V lives in the qk tensor at column v_col0 (one fused q|k|v projection, [B, seq, ld_qk])
int prepare_attn(const void* qk, long long ld_qk, int q_col0, int k_col0, const void* vt, void* out, int ldo, int B,
                 int heads, int seq, int d, float scale, AttnOp* op, int v_col0 = -1);
int run_attn(const AttnOp& op, cudaStream_t st);
const char* attn_last_error();

}  // namespace dxmi
