// ADM / EDM U-Net (models/cm/unet.py:523-790, UNetModel) spec + plan builder.
//   ResBlock._forward        models/cm/unet.py:240-260   (GN32 -> SiLU -> [avgpool | nearest x2] -> conv3 ; emb Linear ;
//                                                          GN32 (FiLM scale/shift) -> SiLU -> conv3 ; + skip)
//   AttentionBlock._forward  models/cm/unet.py:320-332   (GN32 -> qkv 1x1 -> QKVAttentionLegacy :413-441 -> proj 1x1 -> + x)
//   UNetModel.forward        models/cm/unet.py:761-790
// Only the configuration family the DxMI YAMLs use is built: resblock_updown=True, conv_resample unused,
// num_head_channels (or num_heads) fixed per net, dropout 0, no gradient checkpointing.
#include <cmath>
#include <cstdio>

#include "adm_layout.cuh"
#include "attn_tc.cuh"
#include "builder.cuh"

namespace dxmi {

namespace {

void expect(Net& net, const std::string& k, std::vector<int64_t> shape) {
    net.keys.push_back(k);
    net.expect[k] = std::move(shape);
}
void expect_wb(Net& net, const std::string& p, std::vector<int64_t> wshape) {
    const int64_t o = wshape[0];
    expect(net, p + ".weight", std::move(wshape));
    expect(net, p + ".bias", {o});
}

void spec_layers(Net& net, const std::string& prefix, const std::vector<Layer>& layers, int ted) {
    for (size_t j = 0; j < layers.size(); ++j) {
        const Layer& L = layers[j];
        const std::string p = prefix + "." + std::to_string(j);
        if (L.kind == L_CONV) {
            expect_wb(net, p, {L.cout, L.cin, 3, 3});
        } else if (L.kind == L_ATTN) {
            expect_wb(net, p + ".norm", {L.cout});
            expect_wb(net, p + ".qkv", {3 * L.cout, L.cout, 1});   // Conv1d (conv_nd(1, ...)), models/cm/unet.py:294
            expect_wb(net, p + ".proj_out", {L.cout, L.cout, 1});
        } else {
            expect_wb(net, p + ".in_layers.0", {L.cin});
            expect_wb(net, p + ".in_layers.2", {L.cout, L.cin, 3, 3});
            expect_wb(net, p + ".emb_layers.1", {(net.a.use_scale_shift_norm ? 2 : 1) * L.cout, ted});
            expect_wb(net, p + ".out_layers.0", {L.cout});
            expect_wb(net, p + ".out_layers.3", {L.cout, L.cout, 3, 3});
            if (L.cin != L.cout) expect_wb(net, p + ".skip_connection", {L.cout, L.cin, 1, 1});
        }
    }
}

}  // namespace

// state_dict keys in the reference's registration order.
void spec_adm(Net& net) {
    const dxmi_arch_desc& a = net.a;
    const int mc = a.ch, ted = 4 * mc;
    expect_wb(net, "time_embed.0", {ted, mc});
    expect_wb(net, "time_embed.2", {ted, ted});
    if (a.num_classes > 0) expect(net, "label_emb.weight", {a.num_classes, ted});
    BlockList inputs, outputs;
    int mid = 0;
    adm_layout(a, inputs, outputs, &mid);
    for (size_t i = 0; i < inputs.size(); ++i) spec_layers(net, "input_blocks." + std::to_string(i), inputs[i], ted);
    spec_layers(net, "middle_block", {{L_RES, mid, mid}, {L_ATTN, mid, mid}, {L_RES, mid, mid}}, ted);
    for (size_t i = 0; i < outputs.size(); ++i) spec_layers(net, "output_blocks." + std::to_string(i), outputs[i], ted);
    const int ch0 = a.ch_mult[0] * mc;
    expect_wb(net, "out.0", {ch0});
    expect_wb(net, "out.2", {a.out_channels, ch0, 3, 3});
}

struct AdmBuilder : Builder {
    using Builder::Builder;
    float* film = nullptr;
    int film_ld = 0;
    int film_off = 0;
    static constexpr float EPS = 1e-5f;  // models/cm/nn.py:109-116 (GroupNorm32 default eps)

    Act resblock(const std::string& p, Act xa, Act xb, int Cout, LayerKind mode) {
        cur_label = p;
        const dxmi_arch_desc& a = net.a;
        const int H = xa.H, W = xa.W;
        const int Cin = xa.C + xb.C;
        const int Bn = B;
        const bool film_mode = a.use_scale_shift_norm != 0;
        const int emb_cols = (film_mode ? 2 : 1) * Cout;
        bf16* g1 = (bf16*)scratch(0, (size_t)B * H * W * Cin * 2);
        group_norm(xa, xb, p + ".in_layers.0", EPS, 1, nullptr, 0, g1);
        int Ho = H, Wo = W;
        const bf16* conv_in = g1;
        Act xs = xa;  // skip-path input (resampled for up / down blocks)
        const bool up2 = mode == L_UP && up2_ok(Cin, H, W);
        if (mode == L_DOWN || mode == L_UP) {
            if (xb.C) fail("ADM up/down ResBlock with a concatenated input is not a reference configuration");
            Ho = mode == L_DOWN ? H / 2 : H * 2;
            Wo = mode == L_DOWN ? W / 2 : W * 2;
            bf16* gp = (bf16*)scratch(2, (size_t)B * Ho * Wo * Cin * 2);
            bf16* xp = (bf16*)scratch(3, (size_t)B * Ho * Wo * Cin * 2);
            const bf16* xap = xa.p;
            if (mode == L_DOWN) {
                op([=](cudaStream_t st) {
                    avgpool2(g1, gp, Bn, H, W, Cin, ACT_NONE, st);
                    avgpool2(xap, xp, Bn, H, W, Cin, ACT_NONE, st);
                    return (int)cudaGetLastError();
                }, 2);
            } else if (up2) {
                // in_layers conv = four phase convolutions straight from the low-resolution g1 (builder.cuh conv_up2); only the skip path
                // still needs the upsampled x (the residual operand of the out_layers conv)
                op([=](cudaStream_t st) {
                    upsample2x(xap, xp, Bn, H, W, Cin, st);
                    return (int)cudaGetLastError();
                });
            } else {
                op([=](cudaStream_t st) {
                    upsample2x(g1, gp, Bn, H, W, Cin, st);
                    upsample2x(xap, xp, Bn, H, W, Cin, st);
                    return (int)cudaGetLastError();
                }, 2);
            }
            conv_in = gp;
            xs = Act{xp, Cin, Ho, Wo};
        }
        bf16* h1 = (bf16*)scratch(1, (size_t)B * Ho * Wo * Cout * 2);
        StatSpec h1s = stat_spec(Cout, Ho, Wo, true);
        if (up2) {
            h1s.halo = false;
            h1s.P = up2_stats_P(H, W);
            h1s.bytes = (size_t)B * h1s.P * Cout * 2 * sizeof(float);
        }
        float* h1_stats = h1s.P ? (float*)scratch(6, h1s.bytes) : nullptr;
        if (up2) {
            conv_up2(g1, Cin, H, W, p + ".in_layers.2.weight", p + ".in_layers.2.bias", Cout, h1, h1_stats,
                     (!film_mode && film) ? film + film_off : nullptr, film_ld);
        } else {
            dxmi_gemm_desc d = conv_desc(Ho, Wo);
            set_src(d, 0, conv_in, Cin, Cin);
            add_seg(d, 0, 9);
            d.b_ptr = packed_rows(p + ".in_layers.2", {{{p + ".in_layers.2.weight", 0, Cin}}}, nullptr, nullptr);
            d.b_rows = Cout;
            d.b_ld = 9LL * Cin;
            d.bias = f32(p + ".in_layers.2.bias");
            if (!film_mode) {  // h = h + emb_out (models/cm/unet.py:258)
                d.rowvec = film ? film + film_off : nullptr;
                d.ldrv = film_ld;
            }
            d.out = h1;
            d.ldo = Cout;
            d.gn_stats = h1_stats;
            d.gn_halo_P = h1s.halo ? h1s.P : 0;
            gemm(d);
        }
        bf16* g2 = (bf16*)scratch(0, (size_t)B * Ho * Wo * Cout * 2);
        group_norm(Act{h1, Cout, Ho, Wo, h1_stats, h1s.P, h1s.halo, h1s.P > 0}, Act{}, p + ".out_layers.0", EPS, 1, film_mode && film ? film + film_off : nullptr,
                   film_ld, g2);
        film_off += emb_cols;
        Act out = new_act(Cout, Ho, Wo, true, /*conv3x3_s1=*/true);
        {
            dxmi_gemm_desc d = conv_desc(Ho, Wo);
            set_src(d, 0, g2, Cout, Cout);
            add_seg(d, 0, 9);
            long long K = 9LL * Cout;
            if (Cin != Cout) {
                // out_layers conv + 1x1 skip_connection(cat(xa, xb)) accumulated in one TMEM tile
                std::vector<PackPart> parts = {{p + ".out_layers.3.weight", 0, Cout}, {p + ".skip_connection.weight", 0, xs.C}};
                set_src(d, 1, xs.p, xs.C, xs.C);
                add_seg(d, 1, 1);
                K += xs.C;
                if (xb.C) {
                    parts.push_back({p + ".skip_connection.weight", xs.C, xb.C});
                    set_src(d, 2, xb.p, xb.C, xb.C);
                    add_seg(d, 2, 1);
                    K += xb.C;
                }
                d.b_ptr = packed_rows(p + ".out3+skip", {parts}, nullptr, nullptr);
                d.bias = sum_f32(p + ".out3+skip.bias", p + ".out_layers.3.bias", p + ".skip_connection.bias", Cout);
            } else {
                if (xb.C) fail("ADM ResBlock: identity skip over a concatenated input");
                d.b_ptr = packed_rows(p + ".out_layers.3", {{{p + ".out_layers.3.weight", 0, Cout}}}, nullptr, nullptr);
                d.bias = f32(p + ".out_layers.3.bias");
                d.residual = xs.p;
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

    Act attention(const std::string& p, Act x) {
        cur_label = p;
        const dxmi_arch_desc& a = net.a;
        const int C = x.C, H = x.H, W = x.W, HW = H * W;
        const int heads = a.num_head_channels > 0 ? C / a.num_head_channels : a.num_heads;
        const int dh = C / heads;
        const float scale = 1.f / sqrtf((float)dh);  // q and k are each scaled by d^-1/4 (models/cm/unet.py:433)
        const int Bn = B;
        bf16* hn = (bf16*)scratch(0, (size_t)B * HW * C * 2);
        group_norm(x, Act{}, p + ".norm", EPS, 0, nullptr, 0, hn);
        bf16* o = (bf16*)scratch(4, (size_t)B * HW * C * 2);
        const float* qkv_bias = f32(p + ".qkv.bias");
        if (dh == 64 && (HW % 128 == 0 || HW == 64)) {
            // q | k | v = hn . W^T in ONE GEMM (channel layout (three, heads, d): q rows first, then k, then v); the attention kernel
            // reads V [keys][d] straight from it as an MN-major operand - round 1 ran a separate batched V^T GEMM (weights as the A
            // operand) at 0.2-0.3 PFLOP/s: 6 % of the ImageNet-64 forward (profiles/r02_gemm_table_in64_before.txt)
            bf16* qk = (bf16*)scratch(1, (size_t)B * HW * 3 * C * 2);
            {
                dxmi_gemm_desc d = conv_desc(H, W);
                set_src(d, 0, hn, C, C);
                add_seg(d, 0, 1);
                d.b_ptr = packed_rows(p + ".qkv", {{{p + ".qkv.weight", 0, C}}}, nullptr, nullptr);
                d.b_rows = 3 * C;
                d.b_ld = C;
                d.bias = qkv_bias;
                d.out = qk;
                d.ldo = 3 * C;
                gemm(d);
            }
            if (!dry && !err) {
                AttnOp aop;
                int r = prepare_attn(qk, 3LL * C, 0, C, nullptr, o, C, B, heads, HW, dh, scale, &aop, 2 * C);
                if (r) {
                    err = r;
                    engine_set_error("prepare_attn: %s", attn_last_error());
                } else {
                    plan.gemm_flops += aop.flops;
                    {
                        const std::string keep = cur_label;
                        cur_label = "ATTN " + keep;
                        op([aop](cudaStream_t st) {
                            return run_timed_tensor(aop.flops, aop.p.seq, aop.p.seq, 64, 2 * (int)(aop.grid.y * aop.grid.z), st,
                                                    [&] { return run_attn(aop, st); });
                        });
                        cur_label = keep;
                    }
                }
            }
        } else if (HW <= 64) {
            bf16* qkv = (bf16*)scratch(1, (size_t)B * HW * 3 * C * 2);
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, hn, C, C);
            add_seg(d, 0, 1);
            d.b_ptr = packed_rows(p + ".qkv", {{{p + ".qkv.weight", 0, C}}}, nullptr, nullptr);
            d.b_rows = 3 * C;
            d.b_ld = C;
            d.bias = qkv_bias;
            d.out = qkv;
            d.ldo = 3 * C;
            gemm(d);
            op([=](cudaStream_t st) {
                attn_small(qkv, qkv + C, qkv + 2 * C, 3 * C, o, C, Bn, heads, HW, dh, scale, st);
                return (int)cudaGetLastError();
            });
        } else {
            fail("ADM attention: unsupported (sequence length, head dim) combination");
        }
        Act out = new_act(C, H, W);
        {
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

    Act run_layers(const std::string& prefix, const std::vector<Layer>& layers, Act h, Act skip) {
        for (size_t j = 0; j < layers.size(); ++j) {
            const Layer& L = layers[j];
            const std::string p = prefix + "." + std::to_string(j);
            if (L.kind == L_ATTN) {
                h = attention(p, h);
            } else {
                h = resblock(p, h, j == 0 ? skip : Act{}, L.cout, L.kind);
            }
        }
        return h;
    }

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int mc = a.ch, ted = 4 * mc, R = a.resolution;
        Plan* pl = &plan;
        const int Bn = B;
        if (!a.resblock_updown) {
            fail("ADM U-Net: only resblock_updown=True is built (every DxMI EDM config uses it)");
            return;
        }
        plan.eps = (float*)alloc((size_t)B * a.out_channels * R * R * sizeof(float));
        plan.tbuf = (float*)alloc((size_t)B * sizeof(float));
        plan.coef = (float*)alloc((size_t)B * 8 * sizeof(float));

        BlockList inputs, outputs;
        int mid = 0;
        adm_layout(a, inputs, outputs, &mid);
        const std::vector<Layer> middle{{L_RES, mid, mid}, {L_ATTN, mid, mid}, {L_RES, mid, mid}};

        // ---- every ResBlock's emb_layers Linear in execution order -> one batched projection
        std::vector<std::string> wk, bk;
        int TP = 0;
        auto collect = [&](const std::string& prefix, const std::vector<Layer>& layers) {
            for (size_t j = 0; j < layers.size(); ++j)
                if (layers[j].kind == L_RES || layers[j].kind == L_DOWN || layers[j].kind == L_UP) {
                    const std::string p = prefix + "." + std::to_string(j);
                    wk.push_back(p + ".emb_layers.1.weight");
                    bk.push_back(p + ".emb_layers.1.bias");
                    TP += (a.use_scale_shift_norm ? 2 : 1) * layers[j].cout;
                }
        };
        for (size_t i = 1; i < inputs.size(); ++i) collect("input_blocks." + std::to_string(i), inputs[i]);
        collect("middle_block", middle);
        for (size_t i = 0; i < outputs.size(); ++i) collect("output_blocks." + std::to_string(i), outputs[i]);

        float* te = (float*)alloc((size_t)B * mc * 4);
        float* t1 = (float*)alloc((size_t)B * ted * 4);
        float* emb = (float*)alloc((size_t)B * ted * 4);
        film = (float*)alloc((size_t)B * TP * 4);
        film_ld = TP;
        film_off = 0;
        {
            const float* w0 = f32("time_embed.0.weight");
            const float* b0 = f32("time_embed.0.bias");
            const float* w2 = f32("time_embed.2.weight");
            const float* b2 = f32("time_embed.2.bias");
            const float* table = a.num_classes > 0 ? f32("label_emb.weight") : nullptr;
            op([=](cudaStream_t st) {
                // models/cm/unet.py:775-779: emb = time_embed(timestep_embedding(t)) (+ label_emb(y)), all fp32
                timestep_embedding(pl->t, te, Bn, mc, 1, st);
                linear_f32(te, mc, w0, b0, t1, ted, Bn, mc, ted, 0, 0, st);
                linear_f32(t1, ted, w2, b2, emb, ted, Bn, ted, ted, 2, 0, st);
                if (table) {
                    if (!pl->y) return (int)cudaErrorInvalidValue;  // class-conditional net needs labels
                    embedding_add(emb, table, (const long long*)pl->y, Bn, ted, st);
                }
                return (int)cudaGetLastError();
            }, table ? 4 : 3);
            // emb_layers = SiLU -> Linear (cm/unet.py:203-209), every ResBlock at once on the tensor cores
            batched_emb_projection(emb, ted, "emb_layers", wk, bk, film, TP);
        }
        // ---- input conv (x * c_in folded into the load, karras_diffusion.py:349)
        Act h = new_act(inputs[0][0].cout, R, R, /*want_stats=*/true);  // the input conv writes its GroupNorm partials itself
        {
            const float* w = f32("input_blocks.0.0.weight");
            const float* b = f32("input_blocks.0.0.bias");
            bf16* o = h.p;
            float* hst = h.stats;
            const int Cin = a.in_channels, Co = h.C;
            if (Cin != 3 || Co % 32 || Co > 256 || (R * R) % 128 || h.stats_P < R * R / 128 || h.stats_halo)
                fail("ADM input conv: unsupported geometry");
            h.stats_P = R * R / 128;  // conv3x3_first_k publishes one partial per 128-pixel tile
            const bool first_tc = first_tc_option() && conv3x3_first_tc_supported(Cin, R, R, Co);
            op([=](cudaStream_t st) {
                if (first_tc)
                    conv3x3_first_tc(pl->x, pl->x_scale, w, b, o, hst, Bn, R, R, Co, 0, st);
                else
                    conv3x3_first(pl->x, pl->x_scale, w, b, o, hst, Bn, Cin, R, R, Co, 0, st);
                return (int)cudaGetLastError();
            });
        }
        std::vector<Act> hs{h};
        for (size_t i = 1; i < inputs.size(); ++i) {
            h = run_layers("input_blocks." + std::to_string(i), inputs[i], h, Act{});
            hs.push_back(h);
        }
        h = run_layers("middle_block", middle, h, Act{});
        for (size_t i = 0; i < outputs.size(); ++i) {
            Act skip = hs.back();
            hs.pop_back();
            h = run_layers("output_blocks." + std::to_string(i), outputs[i], h, skip);
        }
        // ---- head: GN32 -> SiLU -> conv3 (fp32, models/cm/unet.py:738-742, :789-790)
        bf16* g = (bf16*)scratch(0, (size_t)B * R * R * h.C * 2);
        group_norm(h, Act{}, "out.0", EPS, 1, nullptr, 0, g);
        conv_out_nchw(g, h.C, R, R, "out.2.weight", "out.2.bias", a.out_channels);
    }
};

int build_adm_plan(Net& net, Plan& plan) { return build_two_pass<AdmBuilder>(net, plan); }

}  // namespace dxmi
