// extern "C" surface of libdxmi_b200.so (see include/dxmi_b200.h for the contract of every entry point).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "attn_tc.cuh"
#include "engine.cuh"
#include "kernels_bwd.cuh"
#include "wgrad_tc.cuh"

using namespace dxmi;

struct dxmi_net_s {
    Net net;
};

static thread_local char g_api_err[768] = "";
static void set_err(const char* m) { snprintf(g_api_err, sizeof g_api_err, "%s", m); }

static int g_opt_t_uniform = 1;      // option "t_uniform": DDPM rollout plans evaluate the timestep-embedding path for one row (engine.cuh Plan::t_uniform)
static int g_rollout_split = 1;      // option "rollout_split": sub-batches per rollout (1 = off, the default: measured neutral)
static int g_rollout_split_min = 32; // option "rollout_split_min": smallest sub-batch worth splitting into
static void dxmi_set_rollout_split(int split, int min_sub) {
    if (split >= 1) g_rollout_split = split > 8 ? 8 : split;
    if (min_sub >= 1) g_rollout_split_min = min_sub;
}

extern "C" {

const char* dxmi_last_error(void) { return g_api_err; }
long long dxmi_launch_count(void) { return total_launches(); }

int dxmi_set_option(const char* name, int value) {
    if (!strcmp(name, "block_n_256")) {
        set_block_n_256(value);
        return 0;
    }
    if (!strcmp(name, "halo")) {
        set_halo(value);
        return 0;
    }
    if (!strcmp(name, "gn_unroll")) {
        set_gn_unroll(value);
        return 0;
    }
    if (!strcmp(name, "gn_fused")) {  // read when a plan is built
        set_gn_fused(value);
        return 0;
    }
    if (!strcmp(name, "t_uniform")) {  // DDPM rollouts: one-row timestep-embedding path (read when a rollout fetches its plan)
        g_opt_t_uniform = value;
        return 0;
    }
    if (!strcmp(name, "conv_out_padded")) {  // read when a plan is built
        set_conv_out_padded(value);
        return 0;
    }
    if (!strcmp(name, "s3_m2")) {
        set_s3_m2(value);
        return 0;
    }
    if (!strcmp(name, "s3_stages_max")) {
        set_s3_stages_max(value);
        return 0;
    }
    if (!strcmp(name, "wave_bn")) {  // read when a plan is built
        set_wave_bn(value);
        return 0;
    }
    if (!strcmp(name, "stats16")) {  // read when a plan is built
        set_stats16(value);
        return 0;
    }
    if (!strcmp(name, "shift3")) {  // read when a plan is built
        set_shift3(value);
        return 0;
    }
    if (!strcmp(name, "small_map_bn")) {  // read when a plan is built
        set_small_map_bn(value);
        return 0;
    }
    if (!strcmp(name, "pdl")) {  // read at every launch
        set_pdl(value);
        return 0;
    }
    if (!strcmp(name, "lean_epi")) {  // read when a GEMM is prepared
        set_lean_epi(value);
        return 0;
    }
    if (!strcmp(name, "first_tc")) {  // read when a plan is built
        set_first_tc(value);
        return 0;
    }
    if (!strcmp(name, "up2")) {  // read when a plan is built
        set_up2(value);
        return 0;
    }
    if (!strcmp(name, "attnblk")) {  // read when a plan is built
        set_attnblk(value);
        return 0;
    }
    if (!strcmp(name, "pair_resident_b")) {
        set_pair_resident_b(value);
        return 0;
    }
    if (!strcmp(name, "pair")) {
        set_pair(value);
        return 0;
    }
    if (!strcmp(name, "pair_min")) {
        set_pair_min(value);
        return 0;
    }
    if (!strcmp(name, "gemm_version")) {
        set_gemm_version(value);
        return 0;
    }
    if (!strcmp(name, "dbg_mode")) {
        set_dbg_mode(value);
        return 0;
    }
    if (!strcmp(name, "rollout_split")) {  // read per rollout call
        dxmi_set_rollout_split(value, -1);
        return 0;
    }
    if (!strcmp(name, "rollout_split_min")) {
        dxmi_set_rollout_split(-1, value);
        return 0;
    }
    if (!strcmp(name, "time_gemms")) {
        set_time_gemms(value);
        return 0;
    }
    set_err("unknown option");
    return -1;
}

int dxmi_set_timing_dump(const char* path) {
    set_timing_dump(path);
    return 0;
}

int dxmi_set_debug_buffer(void* dev_ptr) {
    // one buffer, one consumer: DXMI_DBG_ATTNBLK=1 routes it to the fused attention-block kernel (read when a plan is
    // built, tools/attnblk_timeline.py), otherwise to the persistent GEMM kernels (tools/cta_timeline.py)
    if (getenv("DXMI_DBG_ATTNBLK")) set_attnblk_dbg(dev_ptr);
    else set_dbg_times(dev_ptr);
    return 0;
}

int dxmi_gemm_timing(double* ms_total, double* flops_total, long long* launches) {
    return gemm_timing_collect(ms_total, flops_total, launches);
}

int dxmi_aux_timing(int category, double* ms_total, double* bytes_total, long long* launches) {
    return aux_timing_collect(category, ms_total, bytes_total, launches);
}

double dxmi_plan_gemm_flops(dxmi_net_t net, int B) {
    if (!net) return 0;
    auto it = net->net.plans.find(B);
    return it == net->net.plans.end() ? 0.0 : it->second->gemm_flops;
}

int dxmi_create(const dxmi_arch_desc* desc, int device, dxmi_net_t* out) {
    if (!desc || !out) {
        set_err("dxmi_create: null argument");
        return -1;
    }
    dxmi_net_s* h = new dxmi_net_s();
    h->net.a = *desc;
    h->net.device = device;
    switch (desc->arch) {
        case DXMI_ARCH_DDPM_UNET: spec_ddpm(h->net); break;
        case DXMI_ARCH_IGEBM_V2: spec_igebm(h->net); break;
        case DXMI_ARCH_ADM_UNET: spec_adm(h->net); break;
        default:
            delete h;
            set_err("dxmi_create: unknown arch");
            return -2;
    }
    *out = h;
    return 0;
}

void dxmi_destroy(dxmi_net_t net) { delete net; }

int dxmi_num_weights(dxmi_net_t net) { return net ? (int)net->net.keys.size() : -1; }

int dxmi_weight_key(dxmi_net_t net, int i, char* buf, int buflen) {
    if (!net || i < 0 || i >= (int)net->net.keys.size()) return -1;
    snprintf(buf, buflen, "%s", net->net.keys[i].c_str());
    return 0;
}

int dxmi_weight_shape(dxmi_net_t net, int i, int64_t* shape, int* ndim) {
    if (!net || i < 0 || i >= (int)net->net.keys.size()) return -1;
    const auto& s = net->net.expect[net->net.keys[i]];
    *ndim = (int)s.size();
    for (size_t k = 0; k < s.size(); ++k) shape[k] = s[k];
    return 0;
}

int dxmi_bind_weight(dxmi_net_t net, const char* key, const void* dev_ptr, int dtype, const int64_t* shape, int ndim) {
    if (!net || !key || !dev_ptr) {
        set_err("dxmi_bind_weight: null argument");
        return -1;
    }
    auto it = net->net.expect.find(key);
    if (it == net->net.expect.end()) {
        // keys the sampler injects into the net (log_betas, std) are not consumed by the network kernels
        if (!strcmp(key, "log_betas") || !strcmp(key, "std")) return 0;
        snprintf(g_api_err, sizeof g_api_err, "dxmi_bind_weight: unexpected key '%s'", key);
        return -3;
    }
    const auto& es = it->second;
    bool ok = (int)es.size() == ndim;
    for (int i = 0; ok && i < ndim; ++i) ok = es[i] == shape[i];
    if (!ok) {
        snprintf(g_api_err, sizeof g_api_err, "dxmi_bind_weight: shape mismatch for '%s'", key);
        return -4;
    }
    if (dtype != DXMI_F32 && dtype != DXMI_F16) {
        snprintf(g_api_err, sizeof g_api_err, "dxmi_bind_weight: dtype of '%s' must be fp32 or fp16", key);
        return -5;
    }
    Bound& b = net->net.bound[key];
    const bool dtype_changed = b.ptr && b.dtype != dtype;
    // plans capture borrowed fp32 pointers (biases, GroupNorm affine, Linear weights) by value: a moved parameter
    // invalidates them; dxmi_finalize drops and lazily rebuilds the plans
    if (net->net.finalized && b.ptr && b.ptr != dev_ptr) net->net.ptr_moved = true;
    b.ptr = dev_ptr;
    b.dtype = dtype;
    b.shape.assign(shape, shape + ndim);
    if (dtype_changed && net->net.finalized) {
        set_err("dxmi_bind_weight: dtype changed after finalize; create a new handle");
        return -6;
    }
    return 0;
}

static int run_pack(Net& n, cudaStream_t st) {
    for (auto& j : n.pack_jobs) j(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_err(cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// `inst` > 0: further plan instances (own arena) for the same batch size - the sub-batches of a split rollout
static Plan* get_plan(Net& n, int B, int inst = 0, bool t_uniform = false) {
    const int key = B + (inst << 24) + (t_uniform ? (1 << 30) : 0);
    auto it = n.plans.find(key);
    if (it != n.plans.end()) return it->second.get();
    std::unique_ptr<Plan> p(new Plan());
    p->B = B;
    p->t_uniform = t_uniform;
    const size_t jobs_before = n.pack_jobs.size();
    int r = build_plan(n, *p);
    if (r) {
        set_err(engine_last_error());
        if (p->arena) cudaFree(p->arena);
        return nullptr;
    }
    (void)jobs_before;
    Plan* raw = p.get();
    n.plans[key] = std::move(p);
    return raw;
}

int dxmi_finalize(dxmi_net_t net, dxmi_stream_t stream) {
    if (!net) return -1;
    Net& n = net->net;
    for (auto& k : n.keys) {
        auto it = n.bound.find(k);
        if (it == n.bound.end() || !it->second.ptr) {
            snprintf(g_api_err, sizeof g_api_err, "dxmi_finalize: state_dict key '%s' was never bound", k.c_str());
            return -7;
        }
    }
    DeviceGuard guard(n.device);
    if (n.ptr_moved) {
        // a parameter's storage moved (p.data = ..., load_state_dict(assign=True), .to(), EMA swap): every plan's closures
        // hold the old borrowed pointers. Drop the plans (rebuilt lazily per batch size); packed / derived buffers and
        // their pack jobs resolve pointers at run time and stay.
        cudaStreamSynchronize((cudaStream_t)stream);
        n.drop_plans();
        n.ptr_moved = false;
    }
    if (!n.finalized) {
        // building the B=1 plan registers every packed / derived weight and its pack job
        if (!get_plan(n, 1)) return -8;
        // the rollout (uniform-timestep) plan variant packs one more derived weight: register its job before the pack jobs run
        if (n.a.arch == DXMI_ARCH_DDPM_UNET && n.a.precision == 0 && !get_plan(n, 1, 0, true)) return -8;
        n.finalized = true;
    }
    return run_pack(n, (cudaStream_t)stream);
}

int dxmi_repack(dxmi_net_t net, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_repack: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    return run_pack(net->net, (cudaStream_t)stream);
}

size_t dxmi_workspace_bytes(dxmi_net_t net, int B) {
    if (!net) return 0;
    DeviceGuard guard(net->net.device);
    Plan* p = get_plan(net->net, B);
    return p ? p->arena_bytes : 0;
}

static int run_plan(Net& n, Plan* p, cudaStream_t st) {
    static const bool debug_sync = getenv("DXMI_DEBUG_SYNC") != nullptr;
    const char* time_ops = getenv("DXMI_TIME_OPS");  // profiling: CSV path, one CUDA-event-timed row per plan op
    if (time_ops) {
        std::vector<cudaEvent_t> ev(p->ops.size() + 1);
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
        for (size_t i = 0; i < p->ops.size(); ++i) {
            int r = p->ops[i](st);
            if (r) return r;
            cudaEventRecord(ev[i + 1], st);
        }
        cudaEventSynchronize(ev.back());
        FILE* f = fopen(time_ops, "a");
        for (size_t i = 0; i < p->ops.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            if (f) fprintf(f, "%s,%.2f\n", i < p->op_names.size() ? p->op_names[i].c_str() : "?", ms * 1e3);
        }
        if (f) fclose(f);
        for (auto& e : ev) cudaEventDestroy(e);
        count_launches(p->launches_per_run);
        return 0;
    }
    for (size_t i = 0; i < p->ops.size(); ++i) {
        auto& f = p->ops[i];
        int r = f(st);
        if (!r && debug_sync) r = (int)cudaStreamSynchronize(st);
        if (r && debug_sync)
            fprintf(stderr, "dxmi: op %zu/%zu '%s' failed with %d\n", i, p->ops.size(),
                    i < p->op_names.size() ? p->op_names[i].c_str() : "?", r);
        if (r) {
            snprintf(g_api_err, sizeof g_api_err, "kernel launch failed (%d): %s | %s", r,
                     cudaGetErrorString((cudaError_t)r), gemm_op_last_error());
            return r;
        }
    }
    count_launches(p->launches_per_run);
    return 0;
}

int dxmi_unet_forward(dxmi_net_t net, const float* x, const float* x_scale, const float* t, const int64_t* y, float* out,
                      int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_unet_forward: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    if (net->net.a.arch == DXMI_ARCH_IGEBM_V2) {
        set_err("dxmi_unet_forward called on a value-net handle");
        return -2;
    }
    Plan* p = get_plan(net->net, B);
    if (!p) return -3;
    p->x = x;
    p->x_scale = x_scale;
    p->t = t;
    p->y = y;
    p->out = out;
    return run_plan(net->net, p, (cudaStream_t)stream);
}

int dxmi_value_forward(dxmi_net_t net, const float* x, float* out, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_value_forward: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    if (net->net.a.arch != DXMI_ARCH_IGEBM_V2) {
        set_err("dxmi_value_forward called on a U-Net handle");
        return -2;
    }
    Plan* p = get_plan(net->net, B);
    if (!p) return -3;
    p->x = x;
    p->out = out;
    return run_plan(net->net, p, (cudaStream_t)stream);
}

// ---- value-net training (engine_train.cu)
static int run_ops(const std::vector<std::function<int(cudaStream_t)>>& ops, const std::vector<std::string>& names, long long launches,
                   cudaStream_t st) {
    static const bool debug_sync = getenv("DXMI_DEBUG_SYNC") != nullptr;
    for (size_t i = 0; i < ops.size(); ++i) {
        int r = ops[i](st);
        if (!r && debug_sync) r = (int)cudaStreamSynchronize(st);
        if (r) {
            snprintf(g_api_err, sizeof g_api_err, "op %zu/%zu '%s' failed (%d): %s | %s", i, ops.size(),
                     i < names.size() ? names[i].c_str() : "?", r, cudaGetErrorString((cudaError_t)r), gemm_op_last_error());
            return r;
        }
    }
    count_launches(launches);
    return 0;
}

static Plan* get_train_plan(Net& n, int B, cudaStream_t st) {
    auto it = n.train_plans.find(B);
    if (it != n.train_plans.end()) return it->second.get();
    std::unique_ptr<Plan> p(new Plan());
    p->B = B;
    const size_t jobs_before = n.pack_jobs.size();
    int r = build_train_plan(n, *p);
    if (r) {
        set_err(engine_last_error());
        if (p->arena) cudaFree(p->arena);
        return nullptr;
    }
    // the backward introduces new packed (transposed) weights: pack them now
    for (size_t j = jobs_before; j < n.pack_jobs.size(); ++j) n.pack_jobs[j](st);
    Plan* raw = p.get();
    n.train_plans[B] = std::move(p);
    return raw;
}

int dxmi_bind_grad(dxmi_net_t net, const char* key, float* dev_ptr) {
    if (!net || !key) return -1;
    Net& n = net->net;
    if (n.expect.find(key) == n.expect.end()) {
        snprintf(g_api_err, sizeof g_api_err, "dxmi_bind_grad: '%s' is not a state_dict key of this architecture", key);
        return -5;
    }
    n.grad[key] = dev_ptr;
    return 0;
}

int dxmi_value_forward_train(dxmi_net_t net, const float* x, float* out, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_value_forward_train: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    if (net->net.a.arch != DXMI_ARCH_IGEBM_V2) {
        set_err("dxmi_value_forward_train called on a U-Net handle");
        return -2;
    }
    Plan* p = get_train_plan(net->net, B, (cudaStream_t)stream);
    if (!p) return -3;
    p->x = x;
    p->out = out;
    int r = run_ops(p->ops, p->op_names, p->launches_per_run, (cudaStream_t)stream);
    p->fwd_valid = r == 0;
    return r;
}

int dxmi_op_dropout_mask(void* mask_bf16, long long n, float p, unsigned long long seed, unsigned stream_id, dxmi_stream_t stream) {
    dropout_bf16(nullptr, n, p, seed, stream_id, (bf16*)mask_bf16, (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

int dxmi_unet_forward_train(dxmi_net_t net, const float* x, const float* t, float* out, float dropout_p,
                            unsigned long long dropout_seed, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_unet_forward_train: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    if (net->net.a.arch != DXMI_ARCH_DDPM_UNET) {
        set_err("dxmi_unet_forward_train: the U-Net backward is built for the DDPM U-Net only");
        return -2;
    }
    Plan* p = get_train_plan(net->net, B, (cudaStream_t)stream);
    if (!p) return -3;
    p->x = x;
    p->x_scale = nullptr;
    p->t = t;
    p->y = nullptr;
    p->out = out;
    if (dropout_p < 0.f || dropout_p >= 1.f) {
        set_err("dxmi_unet_forward_train: dropout_p must be in [0, 1)");
        return -5;
    }
    p->dropout_p = dropout_p;
    p->dropout_seed = dropout_seed;
    int r = run_ops(p->ops, p->op_names, p->launches_per_run, (cudaStream_t)stream);
    p->fwd_valid = r == 0;
    return r;
}

int dxmi_adm_forward_train(dxmi_net_t net, const float* x, const float* t, const int64_t* y, float* out, float dropout_p,
                           unsigned long long dropout_seed, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_adm_forward_train: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    if (net->net.a.arch != DXMI_ARCH_ADM_UNET) {
        set_err("dxmi_adm_forward_train called on a handle that is not an ADM U-Net");
        return -2;
    }
    if (net->net.a.num_classes > 0 && !y) {
        set_err("dxmi_adm_forward_train: a class-conditional net needs labels");
        return -6;
    }
    if (dropout_p < 0.f || dropout_p >= 1.f) {
        set_err("dxmi_adm_forward_train: dropout_p must be in [0, 1)");
        return -5;
    }
    Plan* p = get_train_plan(net->net, B, (cudaStream_t)stream);
    if (!p) return -3;
    p->x = x;
    p->x_scale = nullptr;
    p->t = t;
    p->y = y;
    p->out = out;
    p->dropout_p = dropout_p;
    p->dropout_seed = dropout_seed;
    int r = run_ops(p->ops, p->op_names, p->launches_per_run, (cudaStream_t)stream);
    p->fwd_valid = r == 0;
    return r;
}

int dxmi_unet_backward(dxmi_net_t net, const float* x, const float* dout, float* dx, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_unet_backward: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    auto it = net->net.train_plans.find(B);
    if (it == net->net.train_plans.end() || !it->second->fwd_valid) {
        set_err("dxmi_unet_backward: no saved activations for this batch size (call dxmi_unet_forward_train first; one backward "
                "per forward)");
        return -4;
    }
    Plan* p = it->second.get();
    p->x = x;
    p->dout = dout;
    p->dx = dx;
    int r = run_ops(p->bwd_ops, p->bwd_names, p->bwd_launches, (cudaStream_t)stream);
    p->fwd_valid = false;
    return r;
}

int dxmi_value_backward(dxmi_net_t net, const float* x, const float* dout, float* dx, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized) {
        set_err("dxmi_value_backward: handle not finalized");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    auto it = net->net.train_plans.find(B);
    if (it == net->net.train_plans.end() || !it->second->fwd_valid) {
        set_err("dxmi_value_backward: no saved activations for this batch size (call dxmi_value_forward_train first; one "
                "backward per forward)");
        return -4;
    }
    Plan* p = it->second.get();
    p->x = x;
    p->dout = dout;
    p->dx = dx;
    int r = run_ops(p->bwd_ops, p->bwd_names, p->bwd_launches, (cudaStream_t)stream);
    p->fwd_valid = false;
    return r;
}

int dxmi_var_step(const float* x, const float* eps, const float* z, const float* a, const float* c, const float* sigma,
                  float* x_next, float* mean, float* control, float* logp, int B, int chw, dxmi_stream_t stream) {
    if (chw % 4) {
        set_err("dxmi_var_step: C*H*W must be a multiple of 4");
        return -1;
    }
    var_step(x, eps, z, a, c, sigma, x_next, mean, control, logp, nullptr, B, chw, (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

int dxmi_edm_step(const float* x, const float* F, const float* z, const float* coef, float* x_next, float* mean, int B,
                  int chw, dxmi_stream_t stream) {
    if (chw % 4) {
        set_err("dxmi_edm_step: C*H*W must be a multiple of 4");
        return -1;
    }
    edm_step(x, F, z, coef, x_next, mean, nullptr, B, chw, (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

// ---- batch-split rollouts -------------------------------------------------------------------------------------------
// Every image's trajectory is independent and the whole path is bitwise batch-invariant, so a rollout of B
// images may run as S independent rollouts of B/S images on S streams: results are bit-identical to the unsplit call.
// What it buys (profiles/r02_split_*): the HBM-bound kernels (GroupNorm apply, transition) and the small-map GEMMs that
// fill 32-128 of the 148 SMs overlap another sub-batch's tensor-bound GEMMs instead of serialising behind them.
struct SubRollout {
    Plan* p;
    cudaStream_t st;
    int b0, Bs;
};

static int split_count(int B) {
    int S = g_rollout_split;
    if (getenv("DXMI_TIME_OPS")) return 1;
    while (S > 1 && (B % S || B / S < g_rollout_split_min)) --S;
    return S < 1 ? 1 : S;
}

static int make_subs(Net& n, int B, cudaStream_t st, std::vector<SubRollout>& subs, bool t_uniform = false) {
    const int S = split_count(B);
    const int Bs = B / S;
    while ((int)n.side_streams.size() < S - 1) {
        cudaStream_t s2;
        cudaEvent_t e2;
        if (cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
            set_err("rollout split: cannot create side stream / event");
            return -30;
        }
        n.side_streams.push_back(s2);
        n.join_events.push_back(e2);
    }
    if (S > 1 && !n.fork_event && cudaEventCreateWithFlags(&n.fork_event, cudaEventDisableTiming) != cudaSuccess) {
        set_err("rollout split: cannot create fork event");
        return -30;
    }
    for (int k = 0; k < S; ++k) {
        Plan* p = get_plan(n, Bs, S > 1 ? k : 0, t_uniform);
        if (!p) return -3;
        subs.push_back({p, k == 0 ? st : n.side_streams[k - 1], k * Bs, Bs});
    }
    return 0;
}
static void fork_subs(Net& n, std::vector<SubRollout>& subs, cudaStream_t st) {
    if (subs.size() < 2) return;
    cudaEventRecord(n.fork_event, st);
    for (size_t k = 1; k < subs.size(); ++k) cudaStreamWaitEvent(subs[k].st, n.fork_event, 0);
}
static void join_subs(Net& n, std::vector<SubRollout>& subs, cudaStream_t st) {
    for (size_t k = 1; k < subs.size(); ++k) {
        cudaEventRecord(n.join_events[k - 1], subs[k].st);
        cudaStreamWaitEvent(st, n.join_events[k - 1], 0);
    }
}
// one network forward per sub-batch, launches interleaved op by op so that an eager (un-graphed) call feeds all streams
static int run_plans(Net& n, std::vector<SubRollout>& subs) {
    if (subs.size() == 1) return run_plan(n, subs[0].p, subs[0].st);
    const size_t nops = subs[0].p->ops.size();
    for (size_t i = 0; i < nops; ++i)
        for (auto& sb : subs) {
            int r = sb.p->ops[i](sb.st);
            if (r) {
                snprintf(g_api_err, sizeof g_api_err, "kernel launch failed (%d): %s | %s", r, cudaGetErrorString((cudaError_t)r),
                         gemm_op_last_error());
                return r;
            }
        }
    for (auto& sb : subs) count_launches(sb.p->launches_per_run);
    return 0;
}

int dxmi_var_rollout(dxmi_net_t net, const float* sched_host, const float* sigma_dev, int T, const float* noise,
                     float* l_sample, float* mean, float* control, float* logp, uint8_t* sample_u8, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized || net->net.a.arch != DXMI_ARCH_DDPM_UNET) {
        set_err("dxmi_var_rollout: needs a finalized DDPM U-Net handle");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    Net& n = net->net;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<SubRollout> subs;
    int r = make_subs(n, B, st, subs, g_opt_t_uniform != 0);
    if (r) return r;
    const long long chw = (long long)n.a.in_channels * n.a.resolution * n.a.resolution;
    const long long bchw = chw * B;
    // x_0 = first noise tensor (var_sampler.py:242)
    cudaMemcpyAsync(l_sample, noise, bchw * sizeof(float), cudaMemcpyDeviceToDevice, st);
    fork_subs(n, subs, st);
    for (int i = 0; i < T; ++i) {
        for (auto& sb : subs) {
            Plan* p = sb.p;
            var_fill(p->tbuf, p->coef, p->coef + sb.Bs, p->coef + 2 * sb.Bs, sb.Bs, sched_host[3 * i + 0], sched_host[3 * i + 1],
                     sched_host[3 * i + 2], sigma_dev + i, sb.st);
            count_launches(1);
            p->x = l_sample + (long long)i * bchw + sb.b0 * chw;
            p->x_scale = nullptr;
            p->t = p->tbuf;
            p->y = nullptr;
            p->out = p->eps;
        }
        r = run_plans(n, subs);
        if (r) return r;
        for (auto& sb : subs) {
            Plan* p = sb.p;
            const long long o = sb.b0 * chw;
            // algorithmic bytes: read x, eps, z; write x' (+ mean, control when requested), fp32
            run_timed_aux(2, 4.0 * sb.Bs * chw * (4 + (mean ? 1 : 0) + (control ? 1 : 0)), sb.st, [&] {
                var_step(p->x, p->eps, noise + (long long)(i + 1) * bchw + o, p->coef, p->coef + sb.Bs, p->coef + 2 * sb.Bs,
                         l_sample + (long long)(i + 1) * bchw + o, mean ? mean + (long long)i * bchw + o : nullptr,
                         control ? control + (long long)i * bchw + o : nullptr, logp ? logp + (long long)i * B + sb.b0 : nullptr,
                         (sample_u8 && i == T - 1) ? sample_u8 + o : nullptr, sb.Bs, (int)chw, sb.st);
            });
            count_launches(1);
        }
    }
    join_subs(n, subs, st);
    return (int)cudaGetLastError();
}

int dxmi_edm_rollout(dxmi_net_t net, const float* sched_host, const float* sigma_noise_dev, int T, const float* noise,
                     const int64_t* y, float* l_sample, float* mean, uint8_t* sample_u8, int B, dxmi_stream_t stream) {
    if (!net || !net->net.finalized || net->net.a.arch != DXMI_ARCH_ADM_UNET) {
        set_err("dxmi_edm_rollout: needs a finalized ADM U-Net handle");
        return -1;
    }
    DeviceGuard guard(net->net.device);
    Net& n = net->net;
    if ((n.a.num_classes > 0) != (y != nullptr)) {
        set_err("dxmi_edm_rollout: must specify y if and only if the model is class-conditional");  // cm/unet.py:770-772
        return -2;
    }
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<SubRollout> subs;
    int r = make_subs(n, B, st, subs);
    if (r) return r;
    const long long chw = (long long)n.a.in_channels * n.a.resolution * n.a.resolution;
    const long long bchw = chw * B;
    cudaMemcpyAsync(l_sample, noise, bchw * sizeof(float), cudaMemcpyDeviceToDevice, st);  // x_0 (already sigma_max * z)
    fork_subs(n, subs, st);
    for (int i = 0; i < T; ++i) {
        for (auto& sb : subs) {
            Plan* p = sb.p;
            float* coef = p->coef;                 // [Bs, 5]
            float* x_scale = p->coef + 5 * sb.Bs;  // [Bs]
            edm_fill(coef, x_scale, p->tbuf, sb.Bs, sched_host + 6 * i, sigma_noise_dev + i, sb.st);
            count_launches(1);
            p->x = l_sample + (long long)i * bchw + sb.b0 * chw;
            p->x_scale = x_scale;
            p->t = p->tbuf;
            p->y = y ? y + sb.b0 : nullptr;
            p->out = p->eps;
        }
        r = run_plans(n, subs);
        if (r) return r;
        for (auto& sb : subs) {
            Plan* p = sb.p;
            const long long o = sb.b0 * chw;
            run_timed_aux(2, 4.0 * sb.Bs * chw * (4 + (mean ? 1 : 0)), sb.st, [&] {
                edm_step(p->x, p->eps, noise + (long long)(i + 1) * bchw + o, p->coef, l_sample + (long long)(i + 1) * bchw + o,
                         mean ? mean + (long long)i * bchw + o : nullptr, (sample_u8 && i == T - 1) ? sample_u8 + o : nullptr, sb.Bs,
                         (int)chw, sb.st);
            });
            count_launches(1);
        }
    }
    join_subs(n, subs, st);
    return (int)cudaGetLastError();
}

int dxmi_quantize_u8(const float* x, uint8_t* out, long long nelem, dxmi_stream_t stream) {
    quantize_u8(x, out, nelem, (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ kernel-level ops

int dxmi_op_conv_gemm(const dxmi_gemm_desc* d, dxmi_stream_t stream) {
    GemmOp op;
    int r = prepare_gemm(*d, &op);
    if (r) {
        set_err(gemm_op_last_error());
        return r;
    }
    r = run_gemm(op, (cudaStream_t)stream);
    if (r) set_err(gemm_op_last_error());
    count_launches(1);
    return r;
}

int dxmi_op_pack_conv_weight(const void* w, int dtype, int Cout, int Cin, int kh, int kw, int c_off, int c_cnt,
                             void* dst_bf16, long long ldk, long long k_off, dxmi_stream_t stream) {
    pack_conv_weight(w, dtype == DXMI_F16, Cout, Cin, kh, kw, c_off, c_cnt, (bf16*)dst_bf16, ldk, k_off,
                     (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

int dxmi_op_halo_tiles_per_image(int H, int W) { return halo_tiles_per_image(H, W); }

int dxmi_op_gn_ws_floats(int N, int HW, int groups) { return N * gn_num_slabs(N, HW) * 2 * groups; }

int dxmi_op_group_norm(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int N, int HW, int groups,
                       float eps, const float* gamma, const float* beta, const float* film, int film_ld, int silu,
                       float* partial_ws, void* out, dxmi_stream_t stream) {
    if (groups > 32 || (C1 + C2) % (8) || C1 % 8 || (C1 + C2) > 2048 || (C1 + C2) % groups) {
        set_err("dxmi_op_group_norm: unsupported channel / group configuration");
        return -1;
    }
    const int slabs = gn_num_slabs(N, HW);
    gn_stats((const bf16*)x1, C1, ld1, (const bf16*)x2, C2, ld2, N, HW, groups, partial_ws, slabs, (cudaStream_t)stream);
    gn_apply((const bf16*)x1, C1, ld1, (const bf16*)x2, C2, ld2, N, HW, groups, eps, gamma, beta, film, film_ld, silu,
             partial_ws, slabs, (bf16*)out, (cudaStream_t)stream);
    count_launches(2);
    return (int)cudaGetLastError();
}

int dxmi_op_attention(const void* qk, long long ld_qk, int q_col0, int k_col0, const void* vt, void* out, int ldo, int B,
                      int heads, int seq, float scale, dxmi_stream_t stream) {
    AttnOp op;
    int r = prepare_attn(qk, ld_qk, q_col0, k_col0, vt, out, ldo, B, heads, seq, 64, scale, &op);
    if (!r) r = run_attn(op, (cudaStream_t)stream);
    if (r) set_err(attn_last_error());
    count_launches(1);
    return r;
}

long long dxmi_op_gn_bwd_ws_floats(int N, int HW, int C) { return gn_bwd_ws_floats(N, HW, C); }
int dxmi_op_group_norm_bwd(const void* x1, int C1, const void* x2, int C2, const void* dy, const float* ab, const float* mr, int N, int HW,
                           int groups, int silu, float* ws, void* dx, float* dgamma, float* dbeta, dxmi_stream_t stream) {
    const int C = C1 + C2;
    if (C % 8 || C1 % 8 || C > 2048 || groups != 32 || C % groups) {
        set_err("dxmi_op_group_norm_bwd: needs 32 groups, channel counts that are multiples of 8, C <= 2048");
        return -1;
    }
    group_norm_bwd((const bf16*)x1, C1, (const bf16*)x2, C2, (const bf16*)dy, ab, mr, N, HW, groups, silu, ws, (bf16*)dx, dgamma, dbeta,
                   (cudaStream_t)stream);
    count_launches(3);
    return (int)cudaGetLastError();
}

int dxmi_op_pack_conv_weight_up2(const void* w, int dtype, int Cout, int Cin, void* dst_bf16, dxmi_stream_t stream) {
    pack_conv_weight_up2(w, dtype == DXMI_F16, Cout, Cin, (bf16*)dst_bf16, (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

int dxmi_op_pack_conv_weight_dgrad(const void* w, int dtype, int Cout, int Cin, int taps, void* dst_bf16, long long ldk,
                                   long long k_off, dxmi_stream_t stream) {
    pack_conv_weight_dgrad(w, dtype == DXMI_F16, Cout, Cin, taps, (bf16*)dst_bf16, ldk, k_off, (cudaStream_t)stream);
    count_launches(1);
    return (int)cudaGetLastError();
}

long long dxmi_op_wgrad_ws_floats(int N, int H, int W, int Cout, int Cin, int taps) {
    WgradOp op;
    // geometry only: the tensor maps are encoded over a dummy (aligned, never dereferenced) base address
    if (prepare_wgrad((const void*)0x1000, (const void*)0x1000, N, H, W, Cout, Cin, taps, &op)) {
        set_err(gemm_last_error());
        return -1;
    }
    return (long long)op.partial_floats;
}

int dxmi_op_conv_wgrad(const void* dy, const void* x, int N, int H, int W, int Cout, int Cin, int taps, float* grad_oihw,
                       int Cin_total, int ci_off, float scale, float* ws, dxmi_stream_t stream) {
    WgradOp op;
    int r = prepare_wgrad(dy, x, N, H, W, Cout, Cin, taps, &op);
    if (!r) r = run_wgrad(op, ws, grad_oihw, Cin_total, ci_off, scale, (cudaStream_t)stream);
    if (r) set_err(gemm_last_error());
    count_launches(2);
    return r;
}

}  // extern "C"

