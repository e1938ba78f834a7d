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

int prepare_attn(const void* qk, long long ld_qk, int q_col0, int k_col0, const void* vt, void* out, int ldo, int B,
                 int heads, int seq, int d, float scale, AttnOp* op);
int run_attn(const AttnOp& op, cudaStream_t st);
const char* attn_last_error();

}  // namespace dxmi
