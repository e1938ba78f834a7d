// Network plans: a Net owns the borrowed state_dict tensors, their packed bf16 copies and one launch plan per
// batch size.  A plan is a flat list of kernel launches (closures) over a bump-allocated activation arena; TMA
// tensor maps are encoded once at plan-build time.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dxmi_b200.h"
#include "gemm_op.cuh"
#include "kernels.cuh"

namespace dxmi {

struct Bound {
    const void* ptr = nullptr;
    int dtype = DXMI_F32;
    std::vector<int64_t> shape;
};

struct Act {  // NHWC bf16 activation
    bf16* p = nullptr;
    int C = 0, H = 0, W = 0;
    float* stats = nullptr;  // GroupNorm partials written by the producing GEMM: [B*H*W/seg][C][2] (null in the dry pass)
    int stats_P = 0;         // partials per image
    bool stats_halo = false; // one partial per halo tile (3x3 stride-1 conv on a 32/64-wide map) instead of per row segment
    bool has_stats = false;  // pass-independent: the sizing (dry) pass must take the same branches as the real one
    int id = -1;             // training plans: pass-independent identity (pointers are null in the sizing pass)
};

struct Plan {
    int B = 0;
    std::vector<std::function<int(cudaStream_t)>> ops;
    std::vector<std::string> op_names;  // debug labels (DXMI_DEBUG_SYNC=1 reports the failing op)
    char* arena = nullptr;
    size_t arena_bytes = 0;
    // per-call I/O (read by the closures at launch time)
    const float* x = nullptr;
    const float* x_scale = nullptr;
    const float* t = nullptr;
    const int64_t* y = nullptr;
    float* out = nullptr;
    // rollout plans of the DDPM sampler: every image of a step has the SAME timestep (var_sampler.py:262-270), so the embedding MLP and
    // the temb_proj row-stack are evaluated for ONE row and every convolution reads that row (row stride 0). Set before the plan is built.
    bool t_uniform = false;
    // training plans (engine_train.cu): backward launch list + its per-call I/O
    std::vector<std::function<int(cudaStream_t)>> bwd_ops;
    std::vector<std::string> bwd_names;
    long long bwd_launches = 0;
    const float* dout = nullptr;  // [B] gradient of the network output
    float* dx = nullptr;          // optional [B,Cin,H,W] fp32 gradient w.r.t. the input
    bool fwd_valid = false;       // a training forward has filled the saved activations
    float dropout_p = 0.f;        // training-mode dropout of this forward/backward pair (DDPM U-Net)
    unsigned long long dropout_seed = 0;
    long long launches_per_run = 0;
    double gemm_flops = 0;
    // rollout scratch (allocated with the plan)
    float* eps = nullptr;   // [B, Cout, H, W]
    float* tbuf = nullptr;  // [B]
    float* coef = nullptr;  // [B, 8]
};

struct Net {
    dxmi_arch_desc a;
    int device = 0;
    std::vector<std::string> keys;
    std::unordered_map<std::string, std::vector<int64_t>> expect;
    std::unordered_map<std::string, Bound> bound;
    // packed / derived weights
    std::unordered_map<std::string, void*> derived;  // name -> device buffer (owned)
    std::vector<std::function<void(cudaStream_t)>> pack_jobs;
    std::vector<void*> owned;
    bool finalized = false;
    bool ptr_moved = false;  // a bound pointer changed after finalize: plans hold stale borrowed fp32 pointers -> rebuilt by dxmi_finalize
    std::map<int, std::unique_ptr<Plan>> plans;
    std::map<int, std::unique_ptr<Plan>> train_plans;      // forward-with-saved-activations + backward, per batch size
    std::unordered_map<std::string, float*> grad;          // dxmi_bind_grad: fp32 gradient buffer per state_dict key (or null)
    // batch-split rollouts (api.cu): sub-batch k > 0 runs its whole T-step chain on side_streams[k-1], concurrently with the
    // others (HBM-bound GroupNorm / transition kernels and under-filled small-map GEMMs of one sub-batch overlap the
    // tensor-bound GEMMs of another); created lazily, before any CUDA-graph capture (GraphedRollout warms up eagerly)
    std::vector<cudaStream_t> side_streams;
    std::vector<cudaEvent_t> join_events;
    cudaEvent_t fork_event = nullptr;
    void drop_plans();  // frees every plan arena (inference + training); plans are rebuilt lazily per batch size
    ~Net();
};

// RAII: make `device` current for the duration of a C-ABI call and restore the caller's device afterwards
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

// spec (expected keys) builders
void spec_ddpm(Net& net);
void spec_igebm(Net& net);
void spec_adm(Net& net);

// plan builders (dry = size-only pass)
int build_plan(Net& net, Plan& plan);
int build_train_plan(Net& net, Plan& plan);       // engine_train.cu (IGEBM value net; dispatches the DDPM U-Net)
int build_unet_train_plan(Net& net, Plan& plan);  // engine_train_unet.cu
int build_adm_train_plan(Net& net, Plan& plan);   // engine_train_adm.cu

void set_gn_fused(int v);
void set_stats16(int v);
void set_conv_out_padded(int v);
void set_up2(int v);
void set_first_tc(int v);
int first_tc_option();
void set_attnblk(int v);
const char* engine_last_error();
void engine_set_error(const char* fmt, ...);
void count_launches(long long n);
long long total_launches();

}  // namespace dxmi
