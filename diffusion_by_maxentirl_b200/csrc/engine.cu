// Plan builders for the three networks on the DxMI sampler path.
//   DDPM U-Net   models/DxMI/unet_small.py:194-332
//   IGEBM V2     models/modules.py:104-163 (+ models/value.py:8-12)
//   ADM U-Net    models/cm/unet.py:523-790            (engine_adm.cu)
#include "engine.cuh"
#include "attn_tc.cuh"
#include "builder.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace dxmi {

static thread_local char g_eng_err[768] = "";
const char* engine_last_error() { return g_eng_err; }
void engine_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_eng_err, sizeof g_eng_err, fmt, ap);
    va_end(ap);
}
static long long g_launches = 0;
void count_launches(long long n) { g_launches += n; }
long long total_launches() { return g_launches; }

static int g_opt_attnblk = 1;  // DDPM AttnBlock at 16x16 as one kernel (attnblk_tc.cu); 0 = the round-1 five-launch form
int attnblk_option() { return g_opt_attnblk; }
void set_attnblk(int v) { g_opt_attnblk = v; }
static int g_opt_stats16 = 0;  // measured (round 2): 16-row direct partials save the 2nd epilogue barrier but cost more in the finalize: -1 % end to end
int stats16_option() { return g_opt_stats16; }
void set_stats16(int v) { g_opt_stats16 = v; }
static int g_opt_first_tc = 1;  // inference plans: first convolution on mma.sync (conv_first.cu) instead of the fp32 FMA kernel
int first_tc_option() { return g_opt_first_tc; }
void set_first_tc(int v) { g_opt_first_tc = v; }
static int g_opt_up2 = 1;  // nearest-2x upsample + 3x3 conv as four 2x2 phase convolutions (builder.cuh conv_up2): 4/9 of the FLOPs, no upsampled tensor
int up2_option() { return g_opt_up2; }
void set_up2(int v) { g_opt_up2 = v; }
static int g_opt_conv_out_padded = 1;  // Cout <= 4 output conv through the persistent kernel's halo mode (builder.cuh conv_out_nchw)
int conv_out_padded_option() { return g_opt_conv_out_padded; }
void set_conv_out_padded(int v) { g_opt_conv_out_padded = v; }
static int g_opt_gn_fused = 0;
int gn_fused_option() { return g_opt_gn_fused; }
void set_gn_fused(int v) { g_opt_gn_fused = v; }

void Net::drop_plans() {
    for (auto& kv : plans)
        if (kv.second && kv.second->arena) cudaFree(kv.second->arena);
    for (auto& kv : train_plans)
        if (kv.second && kv.second->arena) cudaFree(kv.second->arena);
    plans.clear();
    train_plans.clear();
}

Net::~Net() {
    DeviceGuard g(device);
    drop_plans();
    for (void* p : owned) cudaFree(p);
    for (auto s : side_streams) cudaStreamDestroy(s);
    for (auto e : join_events) cudaEventDestroy(e);
    if (fork_event) cudaEventDestroy(fork_event);
}

// ================================================================================================ specs

static void expect_key(Net& net, const std::string& k, std::vector<int64_t> shape) {
    net.keys.push_back(k);
    net.expect[k] = std::move(shape);
}
static void expect_conv(Net& net, const std::string& p, int co, int ci, int k, bool bias = true) {
    expect_key(net, p + ".weight", {co, ci, k, k});
    if (bias) expect_key(net, p + ".bias", {co});
}
static void expect_linear(Net& net, const std::string& p, int o, int i) {
    expect_key(net, p + ".weight", {o, i});
    expect_key(net, p + ".bias", {o});
}
static void expect_norm(Net& net, const std::string& p, int c) {
    expect_key(net, p + ".weight", {c});
    expect_key(net, p + ".bias", {c});
}

static bool has_attn(const dxmi_arch_desc& a, int v) {
    for (int i = 0; i < a.n_attn; ++i)
        if (a.attn_resolutions[i] == v) return true;
    return false;
}

static void spec_ddpm_resblock(Net& net, const std::string& p, int cin, int cout, int temb) {
    expect_norm(net, p + ".norm1", cin);
    expect_conv(net, p + ".conv1", cout, cin, 3);
    expect_linear(net, p + ".temb_proj", cout, temb);
    expect_norm(net, p + ".norm2", cout);
    expect_conv(net, p + ".conv2", cout, cout, 3);
    if (cin != cout) expect_conv(net, p + ".nin_shortcut", cout, cin, 1);
}
static void spec_ddpm_attn(Net& net, const std::string& p, int c) {
    expect_norm(net, p + ".norm", c);
    expect_conv(net, p + ".q", c, c, 1);
    expect_conv(net, p + ".k", c, c, 1);
    expect_conv(net, p + ".v", c, c, 1);
    expect_conv(net, p + ".proj_out", c, c, 1);
}

// Same module registration order as unet_small.Model.__init__ (unet_small.py:195-289) so state_dict order matches.
void spec_ddpm(Net& net) {
    const dxmi_arch_desc& a = net.a;
    const int ch = a.ch, temb = 4 * ch;
    expect_linear(net, "temb.dense.0", temb, ch);
    expect_linear(net, "temb.dense.1", temb, temb);
    expect_conv(net, "conv_in", ch, a.in_channels, 3);
    int res = a.resolution, block_in = ch;
    for (int l = 0; l < a.n_levels; ++l) {
        const int block_out = ch * a.ch_mult[l];
        block_in = ch * (l == 0 ? 1 : a.ch_mult[l - 1]);
        const std::string lp = "down." + std::to_string(l);
        for (int b = 0; b < a.num_res_blocks; ++b) {
            spec_ddpm_resblock(net, lp + ".block." + std::to_string(b), block_in, block_out, temb);
            block_in = block_out;
        }
        if (has_attn(a, res))
            for (int b = 0; b < a.num_res_blocks; ++b) spec_ddpm_attn(net, lp + ".attn." + std::to_string(b), block_out);
        if (l != a.n_levels - 1) {
            expect_conv(net, lp + ".downsample.conv", block_in, block_in, 3);
            res /= 2;
        }
    }
    spec_ddpm_resblock(net, "mid.block_1", block_in, block_in, temb);
    spec_ddpm_attn(net, "mid.attn_1", block_in);
    spec_ddpm_resblock(net, "mid.block_2", block_in, block_in, temb);
    // `up` modules are inserted at the front of the ModuleList while iterating levels in reverse, so the
    // state_dict lists up.0 first; the channel bookkeeping still runs from the deepest level upward.
    std::vector<std::vector<std::pair<std::string, std::vector<int>>>> per_level(a.n_levels);
    {
        int bi = block_in;
        int r = res;
        for (int l = a.n_levels - 1; l >= 0; --l) {
            const int block_out = ch * a.ch_mult[l];
            int skip_in = ch * a.ch_mult[l];
            for (int b = 0; b <= a.num_res_blocks; ++b) {
                if (b == a.num_res_blocks) skip_in = ch * (l == 0 ? 1 : a.ch_mult[l - 1]);
                per_level[l].push_back({"block", {bi + skip_in, block_out}});
                bi = block_out;
            }
            if (has_attn(a, r))
                for (int b = 0; b <= a.num_res_blocks; ++b) per_level[l].push_back({"attn", {block_out}});
            if (l != 0) {
                per_level[l].push_back({"upsample", {bi}});
                r *= 2;
            }
        }
    }
    for (int l = 0; l < a.n_levels; ++l) {
        const std::string lp = "up." + std::to_string(l);
        int bidx = 0, aidx = 0;
        for (auto& e : per_level[l]) {
            if (e.first == "block")
                spec_ddpm_resblock(net, lp + ".block." + std::to_string(bidx++), e.second[0], e.second[1], temb);
            else if (e.first == "attn")
                spec_ddpm_attn(net, lp + ".attn." + std::to_string(aidx++), e.second[0]);
            else
                expect_conv(net, lp + ".upsample.conv", e.second[0], e.second[0], 3);
        }
    }
    expect_norm(net, "norm_out", ch * a.ch_mult[0]);
    expect_conv(net, "conv_out", a.out_channels, ch * a.ch_mult[0], 3);
}

// IGEBMEncoderV2(use_spectral_norm=False, keepdim=False, n_class=None) (modules.py:108-140)
void spec_igebm(Net& net) {
    const int nh = net.a.ch;
    expect_conv(net, "conv1", nh, net.a.in_channels, 3);
    const int cin[6] = {nh, nh, nh, 2 * nh, 2 * nh, 2 * nh};
    const int cout[6] = {nh, nh, 2 * nh, 2 * nh, 2 * nh, 2 * nh};
    const bool down[6] = {true, false, true, false, true, false};
    for (int i = 0; i < 6; ++i) {
        const std::string p = "blocks." + std::to_string(i);
        expect_conv(net, p + ".conv1", cout[i], cin[i], 3);
        expect_conv(net, p + ".conv2", cout[i], cout[i], 3);
        if (cin[i] != cout[i] || down[i]) expect_conv(net, p + ".skip.0", cout[i], cin[i], 1, /*bias=*/false);
    }
    expect_linear(net, "linear", net.a.out_channels, 2 * nh);
    if (net.a.learn_out_scale) expect_linear(net, "out_scale", 1, 1);
}

// ================================================================================================ DDPM U-Net

struct DdpmBuilder : Builder {
    using Builder::Builder;
    float* tproj = nullptr;
    int tproj_ld = 0;
    int tproj_off = 0;
    std::vector<std::string> tproj_w, tproj_b;

    Act resblock(const std::string& p, Act xa, Act xb, int Cout) {
        cur_label = p;
        const int H = xa.H, W = xa.W, HW = H * W;
        const int Cin = xa.C + xb.C;
        bf16* g1 = (bf16*)scratch(0, (size_t)B * HW * Cin * 2);
        group_norm(xa, xb, p + ".norm1", 1e-6f, 1, nullptr, 0, g1);
        bf16* h1 = (bf16*)scratch(1, (size_t)B * HW * Cout * 2);
        const StatSpec h1s = stat_spec(Cout, H, W, true);
        float* h1_stats = h1s.P ? (float*)scratch(6, h1s.bytes) : nullptr;
        {
            long long K;
            int rows;
            bf16* w = packed_rows(p + ".conv1", {{{p + ".conv1.weight", 0, Cin}}}, &K, &rows);
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, g1, Cin, Cin);
            add_seg(d, 0, 9);
            d.b_ptr = w;
            d.b_rows = Cout;
            d.b_ld = 9LL * Cin;
            d.bias = f32(p + ".conv1.bias");
            d.rowvec = tproj ? tproj + tproj_off : nullptr;
            d.ldrv = tproj_ld;
            d.out = h1;
            d.ldo = Cout;
            d.gn_stats = h1_stats;
            d.gn_halo_P = h1s.halo ? h1s.P : 0;
            gemm(d);
        }
        tproj_off += Cout;
        bf16* g2 = (bf16*)scratch(0, (size_t)B * HW * Cout * 2);
        Act h1a{h1, Cout, H, W, h1_stats, h1s.P, h1s.halo, h1s.P > 0};
        group_norm(h1a, Act{}, p + ".norm2", 1e-6f, 1, nullptr, 0, g2);
        Act out = new_act(Cout, H, W, true, /*conv3x3_s1=*/true);
        {
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, g2, Cout, Cout);
            add_seg(d, 0, 9);
            long long K = 9LL * Cout;
            if (Cin != Cout) {
                // conv2(h) + nin_shortcut(cat(xa, xb)) in one accumulator (unet_small.py:128-136)
                std::vector<PackPart> parts = {{p + ".conv2.weight", 0, Cout}, {p + ".nin_shortcut.weight", 0, xa.C}};
                set_src(d, 1, xa.p, xa.C, xa.C);
                add_seg(d, 1, 1);
                K += xa.C;
                if (xb.C) {
                    parts.push_back({p + ".nin_shortcut.weight", xa.C, xb.C});
                    set_src(d, 2, xb.p, xb.C, xb.C);
                    add_seg(d, 2, 1);
                    K += xb.C;
                }
                d.b_ptr = packed_rows(p + ".conv2+nin", {parts}, nullptr, nullptr);
                d.bias = sum_f32(p + ".conv2+nin.bias", p + ".conv2.bias", p + ".nin_shortcut.bias", Cout);
            } else {
                d.b_ptr = packed_rows(p + ".conv2", {{{p + ".conv2.weight", 0, Cout}}}, nullptr, nullptr);
                d.bias = f32(p + ".conv2.bias");
                d.residual = xa.p;
                d.ldr = Cout;
            }
            d.b_rows = Cout;
            d.b_ld = K;
            d.out = out.p;
            d.ldo = Cout;
            want_stats(d, out);
            gemm(d);
        }
        return out;
    }

    Act attn(const std::string& p, Act x) {
        cur_label = p;
        const int C = x.C, H = x.H, W = x.W, HW = H * W;
        if (HW == 256 && C == 256 && x.has_stats && !x.stats_halo && attnblk_option()) {
            // the whole block as one kernel per image pair-cluster (attnblk_tc.cu)
            Act out = new_act(C, H, W);
            if (!out.has_stats || out.stats_P < 2) fail("attnblk: unexpected GroupNorm partial layout of the output");
            out.stats_P = 2;  // the kernel publishes one partial per 128-row half (the buffer may be sized for more)
            bf16* w = packed_rows(p + ".kvqp", {{{p + ".k.weight", 0, C}}, {{p + ".v.weight", 0, C}}, {{p + ".q.weight", 0, C}},
                                                {{p + ".proj_out.weight", 0, C}}}, nullptr, nullptr);
            const float* bias = concat_f32(p + ".kvqp.bias", {p + ".k.bias", p + ".v.bias", p + ".q.bias", p + ".proj_out.bias"});
            const float* gamma = f32(p + ".norm.weight");
            const float* beta = f32(p + ".norm.bias");
            if (!dry && !err) {
                AttnBlkOp aop;
                int r = prepare_attnblk256(x.p, w, bias, gamma, beta, x.stats, x.stats_P, 1e-6f, 1.f / sqrtf((float)C), out.p, out.stats, B, &aop);
                if (r) {
                    err = r;
                    engine_set_error("prepare_attnblk256: %s", gemm_last_error());
                } else {
                    plan.gemm_flops += aop.flops;
                    const std::string keep = cur_label;
                    cur_label = "ATTNBLK " + keep;
                    op([aop](cudaStream_t st) {
                        return run_timed_tensor(aop.flops, 256, 256, 256, 6 * aop.B, st, [&] { return run_attnblk256(aop, st); });
                    });
                    cur_label = keep;
                }
            }
            return out;
        }
        bf16* hn = (bf16*)scratch(0, (size_t)B * HW * C * 2);
        group_norm(x, Act{}, p + ".norm", 1e-6f, 0, nullptr, 0, hn);
        Act out = new_act(C, H, W);
        bf16* o = (bf16*)scratch(4, (size_t)B * HW * C * 2);
        const float scale = 1.f / sqrtf((float)C);
        if (HW <= 64) {
            // q|k|v in one GEMM, then the whole-sequence-in-one-CTA kernel
            bf16* qkv = (bf16*)scratch(1, (size_t)B * HW * 3 * C * 2);
            bf16* w = packed_rows(p + ".qkv",
                                  {{{p + ".q.weight", 0, C}}, {{p + ".k.weight", 0, C}}, {{p + ".v.weight", 0, C}}},
                                  nullptr, nullptr);
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, hn, C, C);
            add_seg(d, 0, 1);
            d.b_ptr = w;
            d.b_rows = 3 * C;
            d.b_ld = C;
            d.bias = concat_f32(p + ".qkv.bias", {p + ".q.bias", p + ".k.bias", p + ".v.bias"});
            d.out = qkv;
            d.ldo = 3 * C;
            gemm(d);
            const int Bn = B;
            op([=](cudaStream_t st) {
                attn_small(qkv, qkv + C, qkv + 2 * C, 3 * C, o, C, Bn, 1, HW, C, scale, st);
                return (int)cudaGetLastError();
            });
        } else if (HW == 128 || HW == 256) {
            bf16* qk = (bf16*)scratch(1, (size_t)B * HW * 2 * C * 2);
            bf16* vT = (bf16*)scratch(2, (size_t)B * HW * C * 2);
            bf16* P = (bf16*)scratch(3, (size_t)B * HW * HW * 2);
            {   // q | k  = hn . [Wq; Wk]^T
                bf16* w = packed_rows(p + ".qk", {{{p + ".q.weight", 0, C}}, {{p + ".k.weight", 0, C}}}, nullptr, nullptr);
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, hn, C, C);
                add_seg(d, 0, 1);
                d.b_ptr = w;
                d.b_rows = 2 * C;
                d.b_ld = C;
                d.bias = concat_f32(p + ".qk.bias", {p + ".q.bias", p + ".k.bias"});
                d.out = qk;
                d.ldo = 2 * C;
                gemm(d);
            }
            {   // V^T[b] = Wv . hn[b]^T  (weights as the A operand, so V lands key-major for the P.V GEMM)
                bf16* w = packed_rows(p + ".v", {{{p + ".v.weight", 0, C}}}, nullptr, nullptr);
                dxmi_gemm_desc d;
                memset(&d, 0, sizeof d);
                d.N = 1;
                d.H = 1;
                d.W = C;  // rows of Wv
                d.out_H = 1;
                d.out_W = C;
                d.stride = 1;
                set_src(d, 0, w, C, C);
                add_seg(d, 0, 1);
                d.b_ptr = hn;
                d.b_rows = HW;
                d.b_ld = C;
                d.b_batch_stride = (long long)HW * C;
                d.batch = B;
                d.b_batched = 1;
                d.bias = f32(p + ".v.bias");
                d.bias_along_m = 1;
                d.out = vT;
                d.ldo = HW;
                d.out_batch_stride = (long long)C * HW;
                d.alpha = 1.f;
                d.rows_per_image = 1;
                gemm(d);
            }
            if (HW == 256 && C == 256) {
                // fused S = q k^T -> softmax -> P v on one CTA per 128 queries (attn256_tc.cu)
                if (!dry && !err) {
                    Attn256Op aop;
                    int r = prepare_attn256(qk, vT, o, C, B, scale, &aop);
                    if (r) {
                        err = r;
                        engine_set_error("prepare_attn256: %s", gemm_last_error());
                    } else {
                        plan.gemm_flops += aop.flops;
                        {
                            const std::string keep = cur_label;
                            cur_label = "ATTN " + keep;
                            op([aop](cudaStream_t st) {
                                return run_timed_tensor(aop.flops, 256, 256, 256, 2 * (int)aop.grid.y, st, [&] { return run_attn256(aop, st); });
                            });
                            cur_label = keep;
                        }
                    }
                }
            } else {
                {   // P = softmax(scale * q k^T)   (row softmax fused in the epilogue; whole row lives in TMEM)
                    dxmi_gemm_desc d;
                    memset(&d, 0, sizeof d);
                    d.N = B;
                    d.H = 1;
                    d.W = HW;
                    d.out_H = 1;
                    d.out_W = HW;
                    d.stride = 1;
                    set_src(d, 0, qk, C, 2 * C);
                    add_seg(d, 0, 1);
                    d.a_batched = 1;
                    d.b_ptr = qk + C;
                    d.b_rows = HW;
                    d.b_ld = 2 * C;
                    d.b_batch_stride = (long long)HW * 2 * C;
                    d.b_batched = 1;
                    d.batch = B;
                    d.alpha = scale;
                    d.softmax = 1;
                    d.out = P;
                    d.ldo = HW;
                    d.out_batch_stride = (long long)HW * HW;
                    d.rows_per_image = 1;
                    gemm(d);
                }
                {   // O = P . V
                    dxmi_gemm_desc d;
                    memset(&d, 0, sizeof d);
                    d.N = B;
                    d.H = 1;
                    d.W = HW;
                    d.out_H = 1;
                    d.out_W = HW;
                    d.stride = 1;
                    set_src(d, 0, P, HW, HW);
                    add_seg(d, 0, 1);
                    d.a_batched = 1;
                    d.b_ptr = vT;
                    d.b_rows = C;
                    d.b_ld = HW;
                    d.b_batch_stride = (long long)C * HW;
                    d.b_batched = 1;
                    d.batch = B;
                    d.alpha = 1.f;
                    d.out = o;
                    d.ldo = C;
                    d.out_batch_stride = (long long)HW * C;
                    d.rows_per_image = 1;
                    gemm(d);
                }
            }
        } else {
            fail("DDPM attention: unsupported sequence length");
        }
        {   // proj_out + residual
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, o, C, C);
            add_seg(d, 0, 1);
            d.b_ptr = packed_rows(p + ".proj_out", {{{p + ".proj_out.weight", 0, C}}}, nullptr, nullptr);
            d.b_rows = C;
            d.b_ld = C;
            d.bias = f32(p + ".proj_out.bias");
            d.residual = x.p;
            d.ldr = C;
            d.out = out.p;
            d.ldo = C;
            want_stats(d, out);
            gemm(d);
        }
        return out;
    }

    Act downsample(const std::string& p, Act x) {
        const int C = x.C;
        Act out = new_act(C, x.H / 2, x.W / 2);
        dxmi_gemm_desc d = conv_desc(x.H, x.W);
        d.out_H = x.H / 2;
        d.out_W = x.W / 2;
        d.stride = 2;
        d.rows_per_image = (x.H / 2) * (x.W / 2);
        set_src(d, 0, x.p, C, C);
        add_seg(d, 0, 9);
        d.b_ptr = packed_rows(p + ".conv", {{{p + ".conv.weight", 0, C}}}, nullptr, nullptr);
        d.b_rows = C;
        d.b_ld = 9LL * C;
        d.bias = f32(p + ".conv.bias");
        d.out = out.p;
        d.ldo = C;
        want_stats(d, out);
        gemm(d);
        return out;
    }

    Act upsample(const std::string& p, Act x) {
        const int C = x.C, H2 = x.H * 2, W2 = x.W * 2;
        if (up2_ok(C, x.H, x.W)) {
            Act out = new_act_up2(C, x.H, x.W);
            conv_up2(x.p, C, x.H, x.W, p + ".conv.weight", p + ".conv.bias", C, out.p, out.stats, nullptr, 0);
            return out;
        }
        bf16* up = (bf16*)scratch(1, (size_t)B * H2 * W2 * C * 2);
        const bf16* xp = x.p;
        const int Bn = B, H = x.H, W = x.W;
        op([=](cudaStream_t st) {
            upsample2x(xp, up, Bn, H, W, C, st);
            return (int)cudaGetLastError();
        });
        Act out = new_act(C, H2, W2, true, /*conv3x3_s1=*/true);
        dxmi_gemm_desc d = conv_desc(H2, W2);
        set_src(d, 0, up, C, C);
        add_seg(d, 0, 9);
        d.b_ptr = packed_rows(p + ".conv", {{{p + ".conv.weight", 0, C}}}, nullptr, nullptr);
        d.b_rows = C;
        d.b_ld = 9LL * C;
        d.bias = f32(p + ".conv.bias");
        d.out = out.p;
        d.ldo = C;
        want_stats(d, out);
        gemm(d);
        return out;
    }

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int ch = a.ch, temb_ch = 4 * ch, R = a.resolution;
        Plan* pl = &plan;
        const int Bn = B;
        // rollout scratch
        plan.eps = (float*)alloc((size_t)B * a.out_channels * R * R * sizeof(float));
        plan.tbuf = (float*)alloc((size_t)B * sizeof(float));
        plan.coef = (float*)alloc((size_t)B * 8 * sizeof(float));

        // ---- collect temb_proj layers in execution order (needed before emitting the single batched projection)
        std::vector<std::string> rb;  // resblock prefixes in execution order
        {
            int res = R;
            for (int l = 0; l < a.n_levels; ++l) {
                for (int b = 0; b < a.num_res_blocks; ++b) rb.push_back("down." + std::to_string(l) + ".block." + std::to_string(b));
                if (l != a.n_levels - 1) res /= 2;
            }
            rb.push_back("mid.block_1");
            rb.push_back("mid.block_2");
            for (int l = a.n_levels - 1; l >= 0; --l)
                for (int b = 0; b <= a.num_res_blocks; ++b) rb.push_back("up." + std::to_string(l) + ".block." + std::to_string(b));
        }
        int TP = 0;
        if (!dry) {
            for (auto& p : rb) {
                const Bound* bw = get(p + ".temb_proj.weight");
                if (!bw) return;
                TP += (int)bw->shape[0];
            }
        } else {
            TP = 1;  // size refined below (dry pass only needs an upper bound; computed exactly from arch)
            int tot = 0;
            for (int l = 0; l < a.n_levels; ++l) tot += a.num_res_blocks * ch * a.ch_mult[l];
            tot += 2 * ch * a.ch_mult[a.n_levels - 1];
            for (int l = a.n_levels - 1; l >= 0; --l) tot += (a.num_res_blocks + 1) * ch * a.ch_mult[l];
            TP = tot;
        }
        // ---- timestep embedding MLP + all temb_proj at once (unet_small.py:296-299, :123)
        // (rollout plans: one timestep per step - one row, read by every image with row stride 0)
        const int Bt = plan.t_uniform ? 1 : B;
        float* te = (float*)alloc((size_t)Bt * ch * 4);
        float* t1 = (float*)alloc((size_t)Bt * temb_ch * 4);
        float* temb = (float*)alloc((size_t)Bt * temb_ch * 4);
        tproj = (float*)alloc((size_t)Bt * TP * 4);
        tproj_ld = plan.t_uniform ? 0 : TP;
        {
            std::vector<std::string> wk, bk;
            for (auto& p : rb) {
                wk.push_back(p + ".temb_proj.weight");
                bk.push_back(p + ".temb_proj.bias");
            }
            const float* w0 = f32("temb.dense.0.weight");
            const float* b0 = f32("temb.dense.0.bias");
            const float* w1 = f32("temb.dense.1.weight");
            const float* b1 = f32("temb.dense.1.bias");
            op([=](cudaStream_t st) {
                timestep_embedding(pl->t, te, Bt, ch, 0, st);
                linear_f32(te, ch, w0, b0, t1, temb_ch, Bt, ch, temb_ch, 0, 0, st);
                linear_f32(t1, temb_ch, w1, b1, temb, temb_ch, Bt, temb_ch, temb_ch, 2, 0, st);
                return (int)cudaGetLastError();
            },
               3);
            // (uniform-t plans: one live row; a SIMT fp32 matrix-vector product is 0.1 ms per step faster still, but then `sample()` and a
            //  loop of `sample_step()` - which cannot assume one timestep per batch - would no longer agree bit for bit)
            batched_emb_projection(temb, temb_ch, "temb_proj", wk, bk, tproj, TP, Bt);
        }
        // ---- conv_in
        Act h0 = new_act(ch, R, R, /*want_stats=*/true);  // conv_in writes its GroupNorm partials itself (one per 128-pixel tile)
        if (a.in_channels != 3 || ch % 32 || ch > 256 || (R * R) % 128 || h0.stats_P < R * R / 128 || h0.stats_halo)
            fail("DDPM conv_in: unsupported geometry");
        h0.stats_P = R * R / 128;  // conv3x3_first_k publishes one partial per 128-pixel tile
        {
            const float* w = f32("conv_in.weight");
            const float* b = f32("conv_in.bias");
            bf16* o = h0.p;
            float* hst = h0.stats;
            const int Cin = a.in_channels;
            const bool first_tc = first_tc_option() && conv3x3_first_tc_supported(Cin, R, R, ch);
            op([=](cudaStream_t st) {
                if (first_tc)
                    conv3x3_first_tc(pl->x, pl->x_scale, w, b, o, hst, Bn, R, R, ch, 0, st);
                else
                    conv3x3_first(pl->x, pl->x_scale, w, b, o, hst, Bn, Cin, R, R, ch, 0, st);
                return (int)cudaGetLastError();
            });
        }
        // ---- down path
        std::vector<Act> hs{h0};
        int res = R;
        for (int l = 0; l < a.n_levels; ++l) {
            const int cout = ch * a.ch_mult[l];
            const std::string lp = "down." + std::to_string(l);
            for (int b = 0; b < a.num_res_blocks; ++b) {
                Act h = resblock(lp + ".block." + std::to_string(b), hs.back(), Act{}, cout);
                if (has_attn(a, res)) h = attn(lp + ".attn." + std::to_string(b), h);
                hs.push_back(h);
            }
            if (l != a.n_levels - 1) {
                hs.push_back(downsample(lp + ".downsample", hs.back()));
                res /= 2;
            }
        }
        // ---- middle
        Act h = hs.back();
        h = resblock("mid.block_1", h, Act{}, h.C);
        h = attn("mid.attn_1", h);
        h = resblock("mid.block_2", h, Act{}, h.C);
        // ---- up path
        for (int l = a.n_levels - 1; l >= 0; --l) {
            const int cout = ch * a.ch_mult[l];
            const std::string lp = "up." + std::to_string(l);
            for (int b = 0; b <= a.num_res_blocks; ++b) {
                Act skip = hs.back();
                hs.pop_back();
                h = resblock(lp + ".block." + std::to_string(b), h, skip, cout);
                if (has_attn(a, res)) h = attn(lp + ".attn." + std::to_string(b), h);
            }
            if (l != 0) {
                h = upsample(lp + ".upsample", h);
                res *= 2;
            }
        }
        // ---- head
        bf16* g = (bf16*)scratch(0, (size_t)B * R * R * h.C * 2);
        group_norm(h, Act{}, "norm_out", 1e-6f, 1, nullptr, 0, g);
        conv_out_nchw(g, h.C, R, R, "conv_out.weight", "conv_out.bias", a.out_channels);
    }
};

// ================================================================================================ IGEBM V2 value net

struct IgebmBuilder : Builder {
    using Builder::Builder;

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int nh = a.ch, R = a.resolution;
        Plan* pl = &plan;
        const int Bn = B;
        if (a.in_channels != 3 || nh % 32 || nh > 256 || (R * R) % 128) fail("IGEBM conv1: unsupported geometry");
        Act h = new_act(nh, R, R, false);
        {
            const float* w = f32("conv1.weight");
            const float* b = f32("conv1.bias");
            bf16* o = h.p;
            const int Cin = a.in_channels;
            const bool first_tc = first_tc_option() && conv3x3_first_tc_supported(Cin, R, R, nh);
            op([=](cudaStream_t st) {
                if (first_tc)
                    conv3x3_first_tc(pl->x, nullptr, w, b, o, nullptr, Bn, R, R, nh, ACT_LRELU02, st);
                else
                    conv3x3_first(pl->x, nullptr, w, b, o, nullptr, Bn, Cin, R, R, nh, ACT_LRELU02, st);
                return (int)cudaGetLastError();
            });
        }
        const int cin[6] = {nh, nh, nh, 2 * nh, 2 * nh, 2 * nh};
        const int cout[6] = {nh, nh, 2 * nh, 2 * nh, 2 * nh, 2 * nh};
        const bool down[6] = {true, false, true, false, true, false};
        for (int i = 0; i < 6; ++i) {
            const std::string p = "blocks." + std::to_string(i);
            const int H = h.H, W = h.W, Ci = cin[i], Co = cout[i];
            const bool has_skip = (Ci != Co) || down[i];
            bf16* h1 = (bf16*)scratch(1, (size_t)B * H * W * Co * 2);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, h.p, Ci, Ci);
                add_seg(d, 0, 9);
                d.b_ptr = packed_rows(p + ".conv1", {{{p + ".conv1.weight", 0, Ci}}}, nullptr, nullptr);
                d.b_rows = Co;
                d.b_ld = 9LL * Ci;
                d.bias = f32(p + ".conv1.bias");
                d.act = ACT_LRELU02;
                d.out = h1;
                d.ldo = Co;
                gemm(d);
            }
            bf16* o = down[i] ? (bf16*)scratch(2, (size_t)B * H * W * Co * 2) : act_alloc(Co, H, W);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, h1, Co, Co);
                add_seg(d, 0, 9);
                long long K = 9LL * Co;
                if (has_skip) {
                    set_src(d, 1, h.p, Ci, Ci);
                    add_seg(d, 1, 1);
                    K += Ci;
                    d.b_ptr = packed_rows(p + ".conv2+skip", {{{p + ".conv2.weight", 0, Co}, {p + ".skip.0.weight", 0, Ci}}},
                                          nullptr, nullptr);
                } else {
                    d.b_ptr = packed_rows(p + ".conv2", {{{p + ".conv2.weight", 0, Co}}}, nullptr, nullptr);
                    d.residual = h.p;
                    d.ldr = Co;
                }
                d.b_rows = Co;
                d.b_ld = K;
                d.bias = f32(p + ".conv2.bias");
                d.act = down[i] ? ACT_NONE : ACT_LRELU02;  // avg_pool comes before the activation (modules.py:96-99)
                d.out = o;
                d.ldo = Co;
                gemm(d);
            }
            if (down[i]) {
                Act nx{act_alloc(Co, H / 2, W / 2), Co, H / 2, W / 2};
                bf16* dst = nx.p;
                op([=](cudaStream_t st) {
                    avgpool2(o, dst, Bn, H, W, Co, ACT_LRELU02, st);
                    return (int)cudaGetLastError();
                });
                h = nx;
            } else {
                h = Act{o, Co, H, W};
            }
        }
        {
            const float* lw = f32("linear.weight");
            const float* lb = f32("linear.bias");
            const float* sw = a.learn_out_scale ? f32("out_scale.weight") : nullptr;
            const float* sb = a.learn_out_scale ? f32("out_scale.bias") : nullptr;
            const bf16* hp = h.p;
            const int HW = h.H * h.W, C = h.C;
            op([=](cudaStream_t st) {
                value_head(hp, Bn, HW, C, lw, lb, sw, sb, pl->out, st);
                return (int)cudaGetLastError();
            });
        }
    }
};

int build_adm_plan(Net& net, Plan& plan);  // engine_adm.cu

int build_plan_f32(Net& net, Plan& plan);  // engine_f32.cu

int build_plan(Net& net, Plan& plan) {
    if (net.a.precision == 1) return build_plan_f32(net, plan);
    if (net.a.precision != 0) {
        engine_set_error("dxmi_arch_desc.precision must be 0 (bf16) or 1 (fp32), got %d", net.a.precision);
        return -25;
    }
    switch (net.a.arch) {
        case DXMI_ARCH_DDPM_UNET: return build_two_pass<DdpmBuilder>(net, plan);
        case DXMI_ARCH_IGEBM_V2: return build_two_pass<IgebmBuilder>(net, plan);
        case DXMI_ARCH_ADM_UNET: return build_adm_plan(net, plan);
        default: engine_set_error("architecture %d not implemented", net.a.arch); return -22;
    }
}

}  // namespace dxmi
