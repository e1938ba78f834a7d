// Trainer-side fused kernels (second half): what the reference's DxMI_Trainer does between the network
// calls of a config-#4 iteration, as a handful of launches instead of hundreds of tiny torch ops.
//   running cost   ((x' - x)^2 / (2 beta)).mean(CHW) forward + backward        trainer.py:163-169 (used :302, :363)
//   clip + Adam    clip_grad_norm_(params, 0.1) + torch.optim.Adam.step()      trainer.py:324-327, :388-389; train_cifar10.py:283-296
// The optimizer kernels are multi-tensor: one device table of (param, grad, exp_avg, exp_avg_sq, numel, lr) rows, work split
// in fixed 64 K-element chunks so that every reduction order depends on the tensor sizes only (deterministic).
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/dxmi_b200.h"

namespace dxmi {

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum in a fixed order; result valid in every thread
__device__ __forceinline__ float block_sum_all(float v, float* sh) {
    v = warp_sum_f(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += sh[i];
    return r;
}

// ---------------------------------------------------------------- running cost
// rc[n] = mean_chw (x'[n] - x[n])^2 / (2 beta[n]);  one CTA per sample, float4 loads
__global__ void running_cost_fwd_k(const float* __restrict__ x, const float* __restrict__ xn, const float* __restrict__ beta,
                                   float* __restrict__ rc, int chw) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    const float4* a = reinterpret_cast<const float4*>(x + (long long)n * chw);
    const float4* b = reinterpret_cast<const float4*>(xn + (long long)n * chw);
    float acc = 0.f;
    for (int i = threadIdx.x; i < chw / 4; i += blockDim.x) {
        const float4 u = a[i], v = b[i];
        const float d0 = v.x - u.x, d1 = v.y - u.y, d2 = v.z - u.z, d3 = v.w - u.w;
        acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float tot = block_sum_all(acc, sh);
    if (threadIdx.x == 0) rc[n] = tot / (2.f * beta[n]) / (float)chw;
}
// d rc[n] / d x'[n] = (x' - x) / (beta[n] * chw);  d / d x = -that.  dx / dxn may be null.
__global__ void running_cost_bwd_k(const float* __restrict__ x, const float* __restrict__ xn, const float* __restrict__ beta,
                                   const float* __restrict__ g, float* __restrict__ dxn, float* __restrict__ dx, int chw) {
    const int n = blockIdx.x;
    const float s = g[n] / (beta[n] * (float)chw);
    const float4* a = reinterpret_cast<const float4*>(x + (long long)n * chw);
    const float4* b = reinterpret_cast<const float4*>(xn + (long long)n * chw);
    for (int i = threadIdx.x; i < chw / 4; i += blockDim.x) {
        const float4 u = a[i], v = b[i];
        const float4 d = make_float4((v.x - u.x) * s, (v.y - u.y) * s, (v.z - u.z) * s, (v.w - u.w) * s);
        if (dxn) reinterpret_cast<float4*>(dxn + (long long)n * chw)[i] = d;
        if (dx) reinterpret_cast<float4*>(dx + (long long)n * chw)[i] = make_float4(-d.x, -d.y, -d.z, -d.w);
    }
}

// ---------------------------------------------------------------- multi-tensor clip + Adam
constexpr int OPT_CHUNK = 65536;  // elements per CTA work item
constexpr int OPT_THREADS = 512;

// chunk -> (tensor, offset): chunk_tensor[k] = row of the table, chunk_first[k] = first element of the chunk inside that tensor
__global__ void __launch_bounds__(OPT_THREADS) sqnorm_chunks_k(const dxmi_opt_tensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                                                               const long long* __restrict__ chunk_first, float* __restrict__ partial) {
    __shared__ float sh[32];
    const int k = blockIdx.x;
    const dxmi_opt_tensor t = tab[chunk_tensor[k]];
    const long long i0 = chunk_first[k];
    const long long n = t.numel - i0 < OPT_CHUNK ? t.numel - i0 : OPT_CHUNK;
    const float* g = t.grad + i0;
    float acc = 0.f;
    if (t.grad) {
        for (long long i = threadIdx.x; i < n; i += OPT_THREADS) {
            const float v = g[i];
            acc = fmaf(v, v, acc);
        }
    }
    const float tot = block_sum_all(acc, sh);
    if (threadIdx.x == 0) partial[k] = tot;
}
// total norm (fixed-order sum of the chunk partials in double precision, like a sequential fp32-safe reference) and the clip
// coefficient of torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)); out = {norm, coef}
__global__ void clip_coef_k(const float* __restrict__ partial, int n_chunks, float max_norm, float* __restrict__ out) {
    __shared__ double shd[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) acc += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += shd[i];
        const float norm = (float)sqrt(tot);
        out[0] = norm;
        float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
        out[1] = coef < 1.f ? coef : 1.f;
    }
}
// torch.optim.Adam.step() (weight_decay = 0, amsgrad = False, maximize = False) on g * coef:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(OPT_THREADS) adam_chunks_k(const dxmi_opt_tensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                                                             const long long* __restrict__ chunk_first, const float* __restrict__ clip,
                                                             float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                                                             int scale_grads_in_place) {
    const int k = blockIdx.x;
    const dxmi_opt_tensor t = tab[chunk_tensor[k]];
    if (!t.grad) return;
    const long long i0 = chunk_first[k];
    const long long n = t.numel - i0 < OPT_CHUNK ? t.numel - i0 : OPT_CHUNK;
    const float coef = clip ? clip[1] : 1.f;
    float* p = t.param + i0;
    float* g = t.grad + i0;
    float* m = t.exp_avg + i0;
    float* v = t.exp_avg_sq + i0;
    const float step_size = t.lr / bc1;
    for (long long i = threadIdx.x; i < n; i += OPT_THREADS) {
        const float gi = g[i] * coef;
        const float mi = m[i] + (1.f - beta1) * (gi - m[i]);  // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        if (scale_grads_in_place) g[i] = gi;
    }
}
__global__ void __launch_bounds__(OPT_THREADS) scale_chunks_k(const dxmi_opt_tensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                                                              const long long* __restrict__ chunk_first, const float* __restrict__ clip) {
    const int k = blockIdx.x;
    const dxmi_opt_tensor t = tab[chunk_tensor[k]];
    if (!t.grad) return;
    const float coef = clip[1];
    if (coef >= 1.f) return;
    const long long i0 = chunk_first[k];
    const long long n = t.numel - i0 < OPT_CHUNK ? t.numel - i0 : OPT_CHUNK;
    float* g = t.grad + i0;
    for (long long i = threadIdx.x; i < n; i += OPT_THREADS) g[i] *= coef;
}

}  // namespace

}  // namespace dxmi

using namespace dxmi;

extern "C" {

int dxmi_running_cost_fwd(const float* state, const float* next_state, const float* beta_next, float* rc, int B, int chw,
                          dxmi_stream_t stream) {
    if (chw % 4) return -1;
    running_cost_fwd_k<<<B, 256, 0, (cudaStream_t)stream>>>(state, next_state, beta_next, rc, chw);
    return (int)cudaGetLastError();
}

int dxmi_running_cost_bwd(const float* state, const float* next_state, const float* beta_next, const float* grad_rc,
                          float* d_next_state, float* d_state, int B, int chw, dxmi_stream_t stream) {
    if (chw % 4) return -1;
    running_cost_bwd_k<<<B, 256, 0, (cudaStream_t)stream>>>(state, next_state, beta_next, grad_rc, d_next_state, d_state, chw);
    return (int)cudaGetLastError();
}

int dxmi_opt_chunk_elems(void) { return OPT_CHUNK; }

int dxmi_opt_grad_norm(const dxmi_opt_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_first_dev, int n_chunks,
                       float max_norm, float* partial_ws, float* norm_coef_out, int scale_grads, dxmi_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_chunks <= 0) return -1;
    sqnorm_chunks_k<<<n_chunks, OPT_THREADS, 0, st>>>(table_dev, chunk_tensor_dev, chunk_first_dev, partial_ws);
    clip_coef_k<<<1, 1024, 0, st>>>(partial_ws, n_chunks, max_norm, norm_coef_out);
    if (scale_grads) scale_chunks_k<<<n_chunks, OPT_THREADS, 0, st>>>(table_dev, chunk_tensor_dev, chunk_first_dev, norm_coef_out);
    return (int)cudaGetLastError();
}

int dxmi_opt_adam_step(const dxmi_opt_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_first_dev, int n_chunks,
                       const float* norm_coef_or_null, float beta1, float beta2, float eps, int step, int scale_grads_in_place,
                       dxmi_stream_t stream) {
    if (n_chunks <= 0 || step < 1) return -1;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_chunks_k<<<n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(table_dev, chunk_tensor_dev, chunk_first_dev, norm_coef_or_null, beta1,
                                                                     beta2, eps, (float)bc1, (float)sqrt(bc2), scale_grads_in_place);
    return (int)cudaGetLastError();
}

}  // extern "C"
