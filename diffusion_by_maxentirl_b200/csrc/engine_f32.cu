// fp32-mode plans (dxmi_arch_desc.precision == 1): the DDPM U-Net (models/DxMI/unet_small.py:292-332) and the IGEBM value
// net (models/modules.py:142-163) on fp32 NHWC activations with CUDA-core kernels (kernels_f32.cu).  Same Net / Plan /
// arena machinery as the bf16 tensor-core plans; one kernel per reference op, no fusion beyond epilogue adds - this mode
// is the "rel-L2 <= 1e-5 per step" leg of the parity contract, not a throughput path.
#include "builder.cuh"
#include "kernels_f32.cuh"

namespace dxmi {

struct ActF {
    float* p = nullptr;
    int C = 0, H = 0, W = 0;
};

struct F32Builder : Builder {
    using Builder::Builder;
    float* falloc(int C, int H, int W) { return (float*)alloc((size_t)B * H * W * C * sizeof(float)); }
    float* fscratch(int slot, int C, int H, int W) { return (float*)scratch(slot, (size_t)B * H * W * C * sizeof(float)); }

    void conv(const std::string& key, ActF x1, ActF x2, int Cout, int k, int stride, float* out, const float* rowvec = nullptr,
              int ldrv = 0, const float* residual = nullptr, int act = 0, bool bias = true) {
        ConvF32 c;
        c.x1 = x1.p;
        c.C1 = x1.C;
        c.x2 = x2.p;
        c.C2 = x2.C;
        c.w = f32(key + ".weight");
        c.bias = bias ? f32(key + ".bias") : nullptr;
        c.rowvec = rowvec;
        c.ldrv = ldrv;
        c.residual = residual;
        c.out = out;
        c.N = B;
        c.H = x1.H;
        c.W = x1.W;
        c.Cout = Cout;
        c.k = k;
        c.stride = stride;
        c.act = act;
        op([c](cudaStream_t st) {
            conv_f32(c, st);
            return (int)cudaGetLastError();
        });
    }
    void gn(const std::string& key, ActF x1, ActF x2, float eps, int silu, float* out) {
        const float* gamma = f32(key + ".weight");
        const float* beta = f32(key + ".bias");
        const int Bn = B, HW = x1.H * x1.W;
        if ((x1.C + x2.C) % 32) fail("fp32 group_norm: channels must be a multiple of 32");
        op([=](cudaStream_t st) {
            group_norm_f32(x1.p, x1.C, x2.p, x2.C, Bn, HW, 32, eps, gamma, beta, silu, out, st);
            return (int)cudaGetLastError();
        });
    }
};

// ================================================================================================ DDPM U-Net
static bool has_attn_f(const dxmi_arch_desc& a, int v) {
    for (int i = 0; i < a.n_attn; ++i)
        if (a.attn_resolutions[i] == v) return true;
    return false;
}

struct DdpmF32Builder : F32Builder {
    using F32Builder::F32Builder;
    float* temb = nullptr;
    int temb_ch = 0;

    // unet_small.py:117-136
    ActF resblock(const std::string& p, ActF xa, ActF xb, int Cout) {
        cur_label = p;
        const int H = xa.H, W = xa.W, Cin = xa.C + xb.C;
        const int Bn = B, tc = temb_ch;
        float* tp = (float*)alloc((size_t)B * Cout * sizeof(float));
        {
            const float* w = f32(p + ".temb_proj.weight");
            const float* b = f32(p + ".temb_proj.bias");
            const float* te = temb;
            op([=](cudaStream_t st) {
                linear_exact_f32(te, tc, w, b, tp, Cout, Bn, tc, Cout, 2, st);
                return (int)cudaGetLastError();
            });
        }
        float* g1 = fscratch(0, Cin, H, W);
        gn(p + ".norm1", xa, xb, 1e-6f, 1, g1);
        float* h1 = fscratch(1, Cout, H, W);
        conv(p + ".conv1", ActF{g1, Cin, H, W}, ActF{}, Cout, 3, 1, h1, tp, Cout);
        float* g2 = fscratch(0, Cout, H, W);
        gn(p + ".norm2", ActF{h1, Cout, H, W}, ActF{}, 1e-6f, 1, g2);
        const float* shortcut = xa.p;
        if (Cin != Cout) {
            float* sc = fscratch(2, Cout, H, W);
            conv(p + ".nin_shortcut", xa, xb, Cout, 1, 1, sc);
            shortcut = sc;
        }
        ActF out{falloc(Cout, H, W), Cout, H, W};
        conv(p + ".conv2", ActF{g2, Cout, H, W}, ActF{}, Cout, 3, 1, out.p, nullptr, 0, shortcut);
        return out;
    }

    // unet_small.py:167-191
    ActF attn(const std::string& p, ActF x) {
        cur_label = p;
        const int C = x.C, H = x.H, W = x.W, HW = H * W;
        if (HW % 8) fail("fp32 attention: sequence length must be a multiple of 8");
        float* hn = fscratch(0, C, H, W);
        gn(p + ".norm", x, ActF{}, 1e-6f, 0, hn);
        float* q = fscratch(1, C, H, W);
        float* k = fscratch(2, C, H, W);
        float* v = fscratch(3, C, H, W);
        float* o = fscratch(4, C, H, W);
        const ActF hna{hn, C, H, W};
        conv(p + ".q", hna, ActF{}, C, 1, 1, q);
        conv(p + ".k", hna, ActF{}, C, 1, 1, k);
        conv(p + ".v", hna, ActF{}, C, 1, 1, v);
        const int Bn = B;
        const float scale = 1.f / sqrtf((float)C);
        op([=](cudaStream_t st) {
            attention_f32(q, k, v, o, Bn, HW, C, scale, st);
            return (int)cudaGetLastError();
        });
        ActF out{falloc(C, H, W), C, H, W};
        conv(p + ".proj_out", ActF{o, C, H, W}, ActF{}, C, 1, 1, out.p, nullptr, 0, x.p);
        return out;
    }

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int ch = a.ch, R = a.resolution;
        temb_ch = 4 * ch;
        Plan* pl = &plan;
        const int Bn = B, tc = temb_ch;
        plan.eps = (float*)alloc((size_t)B * a.out_channels * R * R * sizeof(float));
        plan.tbuf = (float*)alloc((size_t)B * sizeof(float));
        plan.coef = (float*)alloc((size_t)B * 8 * sizeof(float));
        // ---- timestep embedding MLP (unet_small.py:296-299)
        float* te = (float*)alloc((size_t)B * ch * 4);
        float* t1 = (float*)alloc((size_t)B * temb_ch * 4);
        temb = (float*)alloc((size_t)B * temb_ch * 4);
        {
            const float* w0 = f32("temb.dense.0.weight");
            const float* b0 = f32("temb.dense.0.bias");
            const float* w1 = f32("temb.dense.1.weight");
            const float* b1 = f32("temb.dense.1.bias");
            float* tb = temb;
            cur_label = "temb";
            op([=](cudaStream_t st) {
                timestep_embedding(pl->t, te, Bn, ch, 0, st);
                linear_exact_f32(te, ch, w0, b0, t1, tc, Bn, ch, tc, 0, st);
                linear_exact_f32(t1, tc, w1, b1, tb, tc, Bn, tc, tc, 2, st);
                return (int)cudaGetLastError();
            },
               3);
        }
        // ---- conv_in on the NCHW input
        const int Cin = a.in_channels;
        float* xin = fscratch(5, Cin, R, R);
        op([=](cudaStream_t st) {
            nchw_to_nhwc_f32(pl->x, xin, Bn, Cin, R * R, st);
            return (int)cudaGetLastError();
        });
        ActF h0{falloc(ch, R, R), ch, R, R};
        conv("conv_in", ActF{xin, Cin, R, R}, ActF{}, ch, 3, 1, h0.p);
        // ---- down path
        std::vector<ActF> hs{h0};
        int res = R;
        for (int l = 0; l < a.n_levels; ++l) {
            const int cout = ch * a.ch_mult[l];
            const std::string lp = "down." + std::to_string(l);
            for (int b = 0; b < a.num_res_blocks; ++b) {
                ActF h = resblock(lp + ".block." + std::to_string(b), hs.back(), ActF{}, cout);
                if (has_attn_f(a, res)) h = attn(lp + ".attn." + std::to_string(b), h);
                hs.push_back(h);
            }
            if (l != a.n_levels - 1) {
                // Downsample (unet_small.py:69-76): pad right / bottom by one, 3x3 stride-2 conv
                ActF x = hs.back();
                ActF d{falloc(x.C, x.H / 2, x.W / 2), x.C, x.H / 2, x.W / 2};
                conv(lp + ".downsample.conv", x, ActF{}, x.C, 3, 2, d.p);
                hs.push_back(d);
                res /= 2;
            }
        }
        // ---- middle
        ActF h = hs.back();
        h = resblock("mid.block_1", h, ActF{}, h.C);
        h = attn("mid.attn_1", h);
        h = resblock("mid.block_2", h, ActF{}, h.C);
        // ---- up path
        for (int l = a.n_levels - 1; l >= 0; --l) {
            const int cout = ch * a.ch_mult[l];
            const std::string lp = "up." + std::to_string(l);
            for (int b = 0; b <= a.num_res_blocks; ++b) {
                ActF skip = hs.back();
                hs.pop_back();
                h = resblock(lp + ".block." + std::to_string(b), h, skip, cout);
                if (has_attn_f(a, res)) h = attn(lp + ".attn." + std::to_string(b), h);
            }
            if (l != 0) {
                // Upsample (unet_small.py:50-54): nearest 2x, 3x3 conv
                const int C = h.C, H = h.H, W = h.W;
                float* up = fscratch(1, C, 2 * H, 2 * W);
                const float* src = h.p;
                op([=](cudaStream_t st) {
                    upsample2x_f32(src, up, Bn, H, W, C, st);
                    return (int)cudaGetLastError();
                });
                ActF o{falloc(C, 2 * H, 2 * W), C, 2 * H, 2 * W};
                conv(lp + ".upsample.conv", ActF{up, C, 2 * H, 2 * W}, ActF{}, C, 3, 1, o.p);
                h = o;
                res *= 2;
            }
        }
        // ---- head (unet_small.py:329-331)
        float* g = fscratch(0, h.C, R, R);
        gn("norm_out", h, ActF{}, 1e-6f, 1, g);
        const int Co = a.out_channels;
        float* y = fscratch(1, Co, R, R);
        conv("conv_out", ActF{g, h.C, R, R}, ActF{}, Co, 3, 1, y);
        op([=](cudaStream_t st) {
            nhwc_to_nchw_f32(y, pl->out, Bn, Co, R * R, st);
            return (int)cudaGetLastError();
        });
    }
};

// ================================================================================================ IGEBM V2 value net
struct IgebmF32Builder : F32Builder {
    using F32Builder::F32Builder;
    void build() {
        const dxmi_arch_desc& a = net.a;
        const int nh = a.ch, R = a.resolution, Cin = a.in_channels;
        Plan* pl = &plan;
        const int Bn = B;
        float* xin = fscratch(5, Cin, R, R);
        cur_label = "conv1";
        op([=](cudaStream_t st) {
            nchw_to_nhwc_f32(pl->x, xin, Bn, Cin, R * R, st);
            return (int)cudaGetLastError();
        });
        ActF h{falloc(nh, R, R), nh, R, R};
        conv("conv1", ActF{xin, Cin, R, R}, ActF{}, nh, 3, 1, h.p, nullptr, 0, nullptr, /*lrelu*/ 1);
        const int cin[6] = {nh, nh, nh, 2 * nh, 2 * nh, 2 * nh};
        const int cout[6] = {nh, nh, 2 * nh, 2 * nh, 2 * nh, 2 * nh};
        const bool down[6] = {true, false, true, false, true, false};
        for (int i = 0; i < 6; ++i) {
            // ResBlockV2 (modules.py:71-101): conv1 - lrelu - conv2 (+ 1x1 skip | identity) - [avgpool] - lrelu
            const std::string p = "blocks." + std::to_string(i);
            cur_label = p;
            const int H = h.H, W = h.W, Ci = cin[i], Co = cout[i];
            const bool has_skip = (Ci != Co) || down[i];
            float* h1 = fscratch(1, Co, H, W);
            conv(p + ".conv1", h, ActF{}, Co, 3, 1, h1, nullptr, 0, nullptr, 1);
            const float* shortcut = h.p;
            if (has_skip) {
                float* sc = fscratch(2, Co, H, W);
                conv(p + ".skip.0", h, ActF{}, Co, 1, 1, sc, nullptr, 0, nullptr, 0, /*bias=*/false);
                shortcut = sc;
            }
            if (down[i]) {
                float* o = fscratch(3, Co, H, W);
                conv(p + ".conv2", ActF{h1, Co, H, W}, ActF{}, Co, 3, 1, o, nullptr, 0, shortcut, 0);
                ActF nx{falloc(Co, H / 2, W / 2), Co, H / 2, W / 2};
                float* dst = nx.p;
                op([=](cudaStream_t st) {
                    avgpool2_f32(o, dst, Bn, H, W, Co, 1, st);
                    return (int)cudaGetLastError();
                });
                h = nx;
            } else {
                ActF nx{falloc(Co, H, W), Co, H, W};
                conv(p + ".conv2", ActF{h1, Co, H, W}, ActF{}, Co, 3, 1, nx.p, nullptr, 0, shortcut, 1);
                h = nx;
            }
        }
        {
            const float* lw = f32("linear.weight");
            const float* lb = f32("linear.bias");
            const float* sw = a.learn_out_scale ? f32("out_scale.weight") : nullptr;
            const float* sb = a.learn_out_scale ? f32("out_scale.bias") : nullptr;
            const float* hp = h.p;
            const int HW = h.H * h.W, C = h.C;
            cur_label = "head";
            op([=](cudaStream_t st) {
                value_head_f32(hp, Bn, HW, C, lw, lb, sw, sb, pl->out, st);
                return (int)cudaGetLastError();
            });
        }
    }
};

int build_plan_f32(Net& net, Plan& plan) {
    switch (net.a.arch) {
        case DXMI_ARCH_DDPM_UNET: return build_two_pass<DdpmF32Builder>(net, plan);
        case DXMI_ARCH_IGEBM_V2: return build_two_pass<IgebmF32Builder>(net, plan);
        default:
            engine_set_error("fp32 mode covers the DDPM U-Net and the IGEBM value net (the reference's fp32 networks); the ADM "
                             "U-Net runs an fp16 torso in the reference and has no fp32 mode here");
            return -24;
    }
}

}  // namespace dxmi
