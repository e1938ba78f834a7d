// Training plan of the IGEBM value net (trainer.py:244-264 energy update, :276-326 TD updates, :369-389 the
// value term of the sampler loss): a forward pass that keeps every activation the backward needs, and the backward pass
//   head  ->  6 x ResBlockV2 (conv2 / skip wgrad, conv2 dgrad (.) lrelu', conv1 wgrad, conv1 (+ skip) dgrad (.) lrelu')  ->  conv1
// as one static launch list.  Dense parts run on the tensor cores: data gradients are the forward implicit-GEMM kernel on
// transposed, tap-flipped weights (leaky-relu' fused as the epilogue gate), weight gradients the MN-major split-K kernel
// (wgrad_tc.cu).  Since leaky-relu keeps the sign, every lrelu' mask is read from the saved *output* activation.
// Gradients are WRITTEN to the fp32 buffers bound with dxmi_bind_grad (state_dict layouts); unbound keys are skipped.
#include "builder.cuh"
#include "kernels_bwd.cuh"
#include "wgrad_tc.cuh"

namespace dxmi {

struct IgebmTrainBuilder : Builder {
    using Builder::Builder;

    struct Blk {
        Act in;        // block input (activated output of the previous layer)
        bf16* h1;      // lrelu(conv1(in))
        Act out;       // block output (after pool / lrelu)
        int Ci, Co;
        bool down, has_skip;
        std::string p;
    };

    float** gslot(const std::string& key) { return &net.grad[key]; }  // node addresses of unordered_map are stable

    // [rows = Cin of the forward conv][K = sum taps * Cout] data-gradient weights
    bf16* packed_dgrad(const std::string& name, const std::vector<std::string>& keys, int rows, long long* K_out) {
        if (dry) return nullptr;
        long long K = 0;
        for (auto& k : keys) {
            const Bound* b = get(k);
            if (!b) return nullptr;
            K += b->shape[0] * b->shape[2] * b->shape[3];
        }
        if (K_out) *K_out = K;
        bool fresh = false;
        bf16* d = (bf16*)derived_buf("wT:" + name, (size_t)rows * K * sizeof(bf16), &fresh);
        if (!d || !fresh) return d;
        Net* np = &net;
        long long k_off = 0;
        for (auto& k : keys) {
            const Bound* b = get(k);
            const int Cout = (int)b->shape[0], Cin = (int)b->shape[1], taps = (int)(b->shape[2] * b->shape[3]);
            const std::string key = k;
            net.pack_jobs.push_back([np, key, Cout, Cin, taps, d, K, k_off](cudaStream_t st) {
                const Bound& bb = np->bound[key];
                pack_conv_weight_dgrad(bb.ptr, bb.dtype == DXMI_F16, Cout, Cin, taps, d, K, k_off, st);
                count_launches(1);
            });
            k_off += (long long)taps * Cout;
        }
        return d;
    }

    void wgrad(const bf16* dy, const bf16* x, int H, int W, int Cout, int Cin, int taps, const std::string& wkey) {
        if (dry) {
            // worst-case workspace: S * Cout * taps * Cin floats with S <= 148 / base items (prepare_wgrad)
            const int base = (Cout / 128) * taps;
            const size_t S = 148 / base + 1;
            scratch(4, S * Cout * taps * Cin * sizeof(float));
            return;
        }
        if (err) return;
        WgradOp w;
        int r = prepare_wgrad(dy, x, B, H, W, Cout, Cin, taps, &w);
        if (r) {
            err = r;
            engine_set_error("prepare_wgrad(%s): %s", wkey.c_str(), gemm_last_error());
            return;
        }
        float* ws = (float*)scratch(4, w.partial_floats * sizeof(float));
        float** g = gslot(wkey);
        plan.gemm_flops += w.flops;
        op([w, ws, g, Cin](cudaStream_t st) {
            if (!*g) return 0;
            return run_wgrad(w, ws, *g, Cin, 0, 1.f, st);
        }, 2);
    }
    void bias_grad(const bf16* dy, long long rows, int C, const std::string& bkey) {
        float* ws = (float*)scratch(5, (size_t)colsum_ws_floats(rows, C) * sizeof(float));
        float** g = gslot(bkey);
        op([=](cudaStream_t st) {
            if (!*g) return 0;
            colsum_bf16(dy, rows, C, ws, *g, st);
            return (int)cudaGetLastError();
        }, 2);
    }

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int nh = a.ch, R = a.resolution;
        Plan* pl = &plan;
        const int Bn = B;
        if (a.in_channels != 3 || nh % 128 || nh > 256 || (R * R) % 128 || B > 1024)
            fail("IGEBM training plan: needs 3 input channels, nh in {128, 256}, H*W % 128 == 0, batch <= 1024");
        // ------------------------------------------------------------------ forward (activations kept)
        emit_bwd = false;
        Act h = new_act(nh, R, R, false);
        {
            const float* w = f32("conv1.weight");
            const float* b = f32("conv1.bias");
            bf16* o = h.p;
            cur_label = "conv1";
            // (the same kernel choice as the inference plan, engine.cu IgebmBuilder: value(x) is bitwise the same with and without autograd)
            const bool first_tc = first_tc_option() && conv3x3_first_tc_supported(3, R, R, nh);
            op([=](cudaStream_t st) {
                if (first_tc)
                    conv3x3_first_tc(pl->x, nullptr, w, b, o, nullptr, Bn, R, R, nh, ACT_LRELU02, st);
                else
                    conv3x3_first(pl->x, nullptr, w, b, o, nullptr, Bn, 3, R, R, nh, ACT_LRELU02, st);
                return (int)cudaGetLastError();
            });
        }
        const Act h0 = h;
        const int cin[6] = {nh, nh, nh, 2 * nh, 2 * nh, 2 * nh};
        const int cout[6] = {nh, nh, 2 * nh, 2 * nh, 2 * nh, 2 * nh};
        const bool down[6] = {true, false, true, false, true, false};
        std::vector<Blk> blks;
        for (int i = 0; i < 6; ++i) {
            Blk k;
            k.p = "blocks." + std::to_string(i);
            cur_label = k.p;
            const int H = h.H, W = h.W, Ci = cin[i], Co = cout[i];
            k.in = h;
            k.Ci = Ci;
            k.Co = Co;
            k.down = down[i];
            k.has_skip = (Ci != Co) || down[i];
            k.h1 = act_alloc(Co, H, W);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, h.p, Ci, Ci);
                add_seg(d, 0, 9);
                d.b_ptr = packed_rows(k.p + ".conv1", {{{k.p + ".conv1.weight", 0, Ci}}}, nullptr, nullptr);
                d.b_rows = Co;
                d.b_ld = 9LL * Ci;
                d.bias = f32(k.p + ".conv1.bias");
                d.act = ACT_LRELU02;
                d.out = k.h1;
                d.ldo = Co;
                gemm(d);
            }
            bf16* o = down[i] ? (bf16*)scratch(2, (size_t)B * H * W * Co * 2) : act_alloc(Co, H, W);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, k.h1, Co, Co);
                add_seg(d, 0, 9);
                long long K = 9LL * Co;
                if (k.has_skip) {
                    set_src(d, 1, h.p, Ci, Ci);
                    add_seg(d, 1, 1);
                    K += Ci;
                    d.b_ptr = packed_rows(k.p + ".conv2+skip", {{{k.p + ".conv2.weight", 0, Co}, {k.p + ".skip.0.weight", 0, Ci}}},
                                          nullptr, nullptr);
                } else {
                    d.b_ptr = packed_rows(k.p + ".conv2", {{{k.p + ".conv2.weight", 0, Co}}}, nullptr, nullptr);
                    d.residual = h.p;
                    d.ldr = Co;
                }
                d.b_rows = Co;
                d.b_ld = K;
                d.bias = f32(k.p + ".conv2.bias");
                d.act = down[i] ? ACT_NONE : ACT_LRELU02;
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
            k.out = h;
            blks.push_back(k);
        }
        const float* lw = f32("linear.weight");
        const float* lb = f32("linear.bias");
        const float* sw = a.learn_out_scale ? f32("out_scale.weight") : nullptr;
        const float* sb = a.learn_out_scale ? f32("out_scale.bias") : nullptr;
        {
            const bf16* hp = h.p;
            const int HW = h.H * h.W, C = h.C;
            cur_label = "head";
            op([=](cudaStream_t st) {
                value_head(hp, Bn, HW, C, lw, lb, sw, sb, pl->out, st);
                return (int)cudaGetLastError();
            });
        }

        // ------------------------------------------------------------------ backward
        emit_bwd = true;
        cur_label = "bwd head";
        const int C6 = h.C, HW6 = h.H * h.W;
        bf16* dZ = (bf16*)scratch(0, (size_t)B * HW6 * C6 * 2);
        float* S = (float*)alloc((size_t)B * C6 * sizeof(float));
        {
            const bf16* hp = h.p;
            float** g_lw = gslot("linear.weight");
            float** g_lb = gslot("linear.bias");
            float** g_sw = gslot("out_scale.weight");
            float** g_sb = gslot("out_scale.bias");
            const bool has_scale = a.learn_out_scale != 0;
            bf16* dz = dZ;
            op([=](cudaStream_t st) {
                value_head_bwd(hp, pl->dout, lw, sw, dz, S, Bn, HW6, C6, st);
                value_head_param_grads(S, pl->dout, lw, lb, sw, *g_lw, *g_lb, has_scale ? *g_sw : nullptr, has_scale ? *g_sb : nullptr,
                                       Bn, C6, st);
                return (int)cudaGetLastError();
            }, 2);
        }
        int ping = 0;  // dZ lives in scratch slot `ping` (0 / 1)
        for (int i = 5; i >= 0; --i) {
            const Blk& k = blks[i];
            cur_label = "bwd " + k.p;
            const int H = k.in.H, W = k.in.W, Ci = k.Ci, Co = k.Co;
            const long long rows = (long long)B * H * W;
            // grad w.r.t. the conv2 (+ skip) output
            const bf16* d_o = dZ;
            if (k.down) {
                bf16* up = (bf16*)scratch(2, (size_t)rows * Co * 2);
                const bf16* src = dZ;
                op([=](cudaStream_t st) {
                    avgpool2_bwd(src, up, Bn, H, W, Co, st);
                    return (int)cudaGetLastError();
                });
                d_o = up;
            }
            bias_grad(d_o, rows, Co, k.p + ".conv2.bias");
            wgrad(d_o, k.h1, H, W, Co, Co, 9, k.p + ".conv2.weight");
            if (k.has_skip) wgrad(d_o, k.in.p, H, W, Co, Ci, 1, k.p + ".skip.0.weight");
            // dZ1 = conv2^T(d_o) (.) lrelu'(h1)
            bf16* dZ1 = (bf16*)scratch(3, (size_t)rows * Co * 2);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, d_o, Co, Co);
                add_seg(d, 0, 9);
                long long K = 0;
                d.b_ptr = packed_dgrad(k.p + ".conv2", {k.p + ".conv2.weight"}, Co, &K);
                d.b_rows = Co;
                d.b_ld = 9LL * Co;
                d.gate = k.h1;
                d.ldg = Co;
                d.out = dZ1;
                d.ldo = Co;
                gemm(d);
            }
            bias_grad(dZ1, rows, Co, k.p + ".conv1.bias");
            wgrad(dZ1, k.in.p, H, W, Co, Ci, 9, k.p + ".conv1.weight");
            // grad w.r.t. the block input's pre-activation: conv1^T(dZ1) + skip^T(d_o) | + d_o, gated by the input's sign
            bf16* dIn = (bf16*)scratch(1 - ping, (size_t)rows * Ci * 2);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, dZ1, Co, Co);
                add_seg(d, 0, 9);
                long long K = 0;
                if (k.has_skip) {
                    set_src(d, 1, d_o, Co, Co);
                    add_seg(d, 1, 1);
                    d.b_ptr = packed_dgrad(k.p + ".conv1+skip", {k.p + ".conv1.weight", k.p + ".skip.0.weight"}, Ci, &K);
                    d.b_ld = 10LL * Co;
                } else {
                    d.b_ptr = packed_dgrad(k.p + ".conv1", {k.p + ".conv1.weight"}, Ci, &K);
                    d.b_ld = 9LL * Co;
                    d.residual = d_o;
                    d.ldr = Co;
                }
                d.b_rows = Ci;
                d.gate = k.in.p;
                d.ldg = Ci;
                d.out = dIn;
                d.ldo = Ci;
                gemm(d);
            }
            dZ = dIn;
            ping = 1 - ping;
        }
        // ---- first convolution: dZ is the gradient w.r.t. conv1's pre-activation
        cur_label = "bwd conv1";
        bias_grad(dZ, (long long)B * R * R, nh, "conv1.bias");
        {
            float* ws = (float*)scratch(4, (size_t)B * (R / 4) * nh * 27 * sizeof(float));
            float** g = gslot("conv1.weight");
            const bf16* dz = dZ;
            op([=](cudaStream_t st) {
                if (!*g) return 0;
                conv_first_wgrad(dz, pl->x, ws, *g, Bn, R, R, nh, st);
                return (int)cudaGetLastError();
            }, 2);
        }
        {
            // dx (only when the caller asks for it: sampler update / guidance): 3-row data-gradient GEMM, fp32 NCHW output
            dxmi_gemm_desc d = conv_desc(R, R);
            set_src(d, 0, dZ, nh, nh);
            add_seg(d, 0, 9);
            long long K = 0;
            d.b_ptr = packed_dgrad("conv1", {"conv1.weight"}, 3, &K);
            d.b_rows = 3;
            d.b_ld = 9LL * nh;
            d.out = (void*)16;  // patched per call
            d.ldo = 3;
            d.out_fp32 = 1;
            d.out_nchw = 1;
            d.block_n = 32;
            if (!dry && !err) {
                GemmOp g;
                int r = prepare_gemm(d, &g);
                if (r) {
                    err = r;
                    engine_set_error("prepare_gemm(dx): %s", gemm_op_last_error());
                } else if (g.use_v2) {
                    fail("internal: dx GEMM must use the direct-store kernel");
                } else {
                    op([g, pl](cudaStream_t st) {
                        if (!pl->dx) return 0;
                        GemmOp g2 = g;
                        g2.p.out = pl->dx;
                        return run_gemm(g2, st);
                    });
                }
            }
        }
        emit_bwd = false;
    }
};

int build_train_plan(Net& net, Plan& plan) {
    if (net.a.arch == DXMI_ARCH_DDPM_UNET) return build_unet_train_plan(net, plan);
    if (net.a.arch == DXMI_ARCH_ADM_UNET) return build_adm_train_plan(net, plan);
    if (net.a.arch != DXMI_ARCH_IGEBM_V2) {
        engine_set_error("training plans exist for the IGEBM value net, the DDPM U-Net and the ADM U-Net");
        return -26;
    }
    if (net.a.precision != 0) {
        engine_set_error("training plans run in bf16 mode only");
        return -27;
    }
    return build_two_pass<IgebmTrainBuilder>(net, plan);
}

}  // namespace dxmi
