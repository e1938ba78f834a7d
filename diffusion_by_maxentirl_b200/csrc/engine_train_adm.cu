// Training plan of the ADM / EDM U-Net (models/cm/unet.py:523-790 under autograd; SURVEY 8f rank 4: the EDM sampler update,
// trainer.py:693-746 update_sampler_mixed_precision -> one `sampler.sample_step` = one U-Net evaluation per optimizer step,
// train_image_large.py:155-169,259-263).  Same reverse-mode structure as the DDPM plan (engine_train_unet.cu, train_builder.cuh):
// a forward that keeps what the backward needs and a backward launch list from d loss / d F to every parameter gradient.
//   ResBlock    : GN32 -> SiLU -> [avgpool | nearest x2] -> conv3 ; GN32 with FiLM (1 + scale, shift from emb_layers) -> SiLU ->
//                 [dropout] -> conv3 ; + skip (identity / 1x1, on the resampled input for up / down blocks)  (cm/unet.py:240-260)
//                 backward: FiLM GroupNorm backward = group_norm_bwd on the per-image effective gamma + gn_bwd_film_params
//                 (d scale, d shift -> the emb_layers Linear stack; gamma / beta weighted by 1 + scale); avg-pool / nearest
//                 backward are each other's forward kernels up to a factor
//   Attention   : GN32 -> qkv 1x1 -> heads of d = 64 -> proj 1x1 -> + x (cm/unet.py:320-332, :413-441).  Forward: the fused
//                 flash kernel (attn_tc.cu).  Backward per head, materialised: S = scale q k^T and dP = dO v^T as batched tcgen05
//                 GEMMs straight from the fused qkv tensor (head = column offset), softmax / softmax-backward row kernels,
//                 dV = P^T dO, dQ = dS K, dK = dS^T Q on explicitly transposed operands; seq <= 64: one SIMT CTA per (image, head)
//   emb path    : d film [B, sum 2 Cout] -> stacked emb_layers backward -> time_embed.2 / .0, label_emb rows (deterministic)
// Not built: use_scale_shift_norm = False with dropout, conv_resample (no DxMI EDM config uses them).
#include <cmath>

#include "adm_layout.cuh"
#include "attn_tc.cuh"
#include "train_builder.cuh"

namespace dxmi {

struct AdmTrainBuilder : TrainBuilder {
    using TrainBuilder::TrainBuilder;

    struct Rec {
        int kind = 0;  // 0 conv_in, 1 resblock, 2 attn, 5 head
        std::string p;
        Act xa, xb, out;
        int Cout = 0, mode = L_RES, Ho = 0, Wo = 0;
        bf16 *g1 = nullptr, *gp = nullptr, *xp = nullptr, *h1 = nullptr, *g2 = nullptr;
        GnSave n1, n2;
        int film_off = 0;
        unsigned drop_stream = 0;
        // attention
        bf16 *hn = nullptr, *qkv = nullptr, *o = nullptr;
        int heads = 0;
    };
    std::vector<Rec> tape;
    float* film = nullptr;
    int TP = 0, ted = 0;
    bool film_mode = true;

    // weight gradient with any Cout that is a multiple of 64 (the tcgen05 kernel works on 128-row slices: a trailing 64 rows are
    // covered by an overlapping slice that recomputes - identically - 64 rows of the previous one)
    void wgrad_any(const bf16* dy, long long dy_ld, const bf16* x, long long x_ld, int H, int W, int Cout, int Cin, int taps,
                   const std::string& wkey, int Cin_total, int ci_off) {
        if (Cout % 128 == 0) {
            wgrad_sliced(dy, dy_ld, x, x_ld, H, W, Cout, Cin, taps, wkey, Cin_total, ci_off);
            return;
        }
        if (Cout < 128 || Cout % 64) {
            fail("ADM training: convolution widths must be multiples of 64 and at least 128");
            return;
        }
        for (int co = 0; co < Cout;) {
            const int c0 = co + 128 <= Cout ? co : Cout - 128;
            for (int ci = 0; ci < Cin;) {
                int c = Cin - ci;
                if (c > 256) c = 256;
                wgrad_rows(dy + c0, dy_ld, x + ci, x_ld, H, W, 128, c, taps, wkey, Cin_total, ci_off + ci, c0);
                ci += c;
            }
            co = c0 + 128;
        }
    }
    // TrainBuilder::wgrad with a destination row offset
    void wgrad_rows(const bf16* dy, long long dy_ld, const bf16* x, long long x_ld, int H, int W, int Cout, int Cin, int taps,
                    const std::string& wkey, int Cin_total, int ci_off, int row_off) {
        const int base = (Cout / 128) * taps;
        if (dry) {
            scratch(4, (size_t)(148 / (base > 0 ? base : 1) + 1) * Cout * taps * Cin * sizeof(float));
            return;
        }
        if (err) return;
        WgradOp w;
        int r = prepare_wgrad(dy, x, B, H, W, Cout, Cin, taps, &w, dy_ld, x_ld);
        if (r) {
            err = r;
            engine_set_error("prepare_wgrad(%s): %s", wkey.c_str(), gemm_last_error());
            return;
        }
        float* ws = (float*)scratch(4, w.partial_floats * sizeof(float));
        float** g = gslot(wkey);
        plan.gemm_flops += w.flops;
        const long long goff = (long long)row_off * Cin_total * taps;
        op([w, ws, g, Cin_total, ci_off, goff](cudaStream_t st) {
            if (!*g) return 0;
            return run_wgrad(w, ws, *g + goff, Cin_total, ci_off, 1.f, st);
        }, 2);
    }

    // ---------------------------------------------------------------- forward
    Act resblock(const std::string& p, Act xa, Act xb, int Cout, int mode, int film_off) {
        cur_label = p;
        Rec r;
        r.kind = 1;
        r.p = p;
        r.xa = xa;
        r.xb = xb;
        r.Cout = Cout;
        r.mode = mode;
        r.film_off = film_off;
        const int H = xa.H, W = xa.W, Cin = xa.C + xb.C, Bn = B;
        r.g1 = act_alloc(Cin, H, W);
        r.n1 = gn_fwd(xa, xb, p + ".in_layers.0", 1, r.g1);
        int Ho = H, Wo = W;
        const bf16* conv_src = r.g1;
        Act xs = xa;
        if (mode == L_DOWN || mode == L_UP) {
            if (xb.C) fail("ADM up/down ResBlock with a concatenated input is not a reference configuration");
            Ho = mode == L_DOWN ? H / 2 : H * 2;
            Wo = mode == L_DOWN ? W / 2 : W * 2;
            r.gp = act_alloc(Cin, Ho, Wo);
            r.xp = act_alloc(Cin, Ho, Wo);
            bf16 *g1 = r.g1, *gp = r.gp, *xp = r.xp;
            const bf16* xap = xa.p;
            if (mode == L_DOWN) {
                op([=](cudaStream_t st) {
                    avgpool2(g1, gp, Bn, H, W, Cin, ACT_NONE, st);
                    avgpool2(xap, xp, Bn, H, W, Cin, ACT_NONE, st);
                    return (int)cudaGetLastError();
                }, 2);
            } else {
                op([=](cudaStream_t st) {
                    upsample2x(g1, gp, Bn, H, W, Cin, st);
                    upsample2x(xap, xp, Bn, H, W, Cin, st);
                    return (int)cudaGetLastError();
                }, 2);
            }
            conv_src = r.gp;
            xs = Act{r.xp, Cin, Ho, Wo};
        }
        r.Ho = Ho;
        r.Wo = Wo;
        // h = in_layers conv (+ emb_out when the block is not FiLM-conditioned, cm/unet.py:258)
        Act h1 = conv3(p + ".in_layers.2", conv_src, Cin, Ho, Wo, Cout, film_mode ? nullptr : film + film_off, TP, nullptr);
        r.h1 = h1.p;
        r.g2 = act_alloc(Cout, Ho, Wo);
        r.n2 = gn_fwd(h1, Act{}, p + ".out_layers.0", 1, r.g2, film_mode ? film + film_off : nullptr, TP);
        {
            Plan* pl = &plan;
            bf16* g2 = r.g2;
            const long long n = (long long)B * Ho * Wo * Cout;
            const unsigned sid = (unsigned)tape.size();
            r.drop_stream = sid;
            op([=](cudaStream_t st) {
                if (pl->dropout_p > 0.f) dropout_bf16(g2, n, pl->dropout_p, pl->dropout_seed, sid, nullptr, st);
                return (int)cudaGetLastError();
            });
        }
        Act out = mk(Cout, Ho, Wo);
        {
            dxmi_gemm_desc d = conv_desc(Ho, Wo);
            set_src(d, 0, r.g2, Cout, Cout);
            add_seg(d, 0, 9);
            long long K = 9LL * Cout;
            if (Cin != Cout) {
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
        r.out = out;
        tape.push_back(r);
        return out;
    }

    Act attention(const std::string& p, Act x) {
        cur_label = p;
        const dxmi_arch_desc& a = net.a;
        Rec r;
        r.kind = 2;
        r.p = p;
        r.xa = x;
        const int C = x.C, H = x.H, W = x.W, S = H * W, Bn = B;
        const int heads = a.num_head_channels > 0 ? C / a.num_head_channels : a.num_heads;
        const int dh = C / heads;
        const float scale = 1.f / sqrtf((float)dh);
        r.heads = heads;
        r.hn = act_alloc(C, H, W);
        r.n1 = gn_fwd(x, Act{}, p + ".norm", 0, r.hn);
        r.qkv = (bf16*)alloc((size_t)B * S * 3 * C * sizeof(bf16));
        {
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, r.hn, C, C);
            add_seg(d, 0, 1);
            d.b_ptr = packed_rows(p + ".qkv", {{{p + ".qkv.weight", 0, C}}}, nullptr, nullptr);
            d.b_rows = 3 * C;
            d.b_ld = C;
            d.bias = f32(p + ".qkv.bias");
            d.out = r.qkv;
            d.ldo = 3 * C;
            gemm(d);
        }
        r.o = act_alloc(C, H, W);
        bf16 *qkv = r.qkv, *o = r.o;
        if (dh != 64) fail("ADM training attention: head dimension must be 64");
        if (S % 128 == 0 || S == 64) {
            if (!dry && !err) {
                AttnOp aop;
                int rr = prepare_attn(qkv, 3LL * C, 0, C, nullptr, o, C, B, heads, S, dh, scale, &aop, 2 * C);
                if (rr) {
                    err = rr;
                    engine_set_error("prepare_attn: %s", attn_last_error());
                } else {
                    plan.gemm_flops += aop.flops;
                    op([aop](cudaStream_t st) { return run_attn(aop, st); });
                }
            }
        } else if (S < 64) {
            op([=](cudaStream_t st) {
                attn_small(qkv, qkv + C, qkv + 2 * C, 3 * C, o, C, Bn, heads, S, dh, scale, st);
                return (int)cudaGetLastError();
            });
        } else {
            fail("ADM training attention: unsupported sequence length");
        }
        Act out = mk(C, H, W);
        {
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, r.o, C, C);
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
        r.out = out;
        tape.push_back(r);
        return out;
    }

    // ---------------------------------------------------------------- backward
    void gn_bwd_film(const std::string& pfx, Act x, const bf16* dy, const GnSave& s, bf16* dx, const float* film_p, float* d_film_p) {
        const int C = x.C, HW = x.H * x.W, Bn = B, tp = TP;
        float* ws = (float*)scratch(6, (size_t)gn_bwd_ws_floats(B, HW, C) * sizeof(float));
        float** gg = gslot(pfx + ".weight");
        float** gb = gslot(pfx + ".bias");
        const float* gamma = f32(pfx + ".weight");
        const float* beta = f32(pfx + ".bias");
        const bf16* xp = x.p;
        const float *ab = s.ab, *mr = s.mr;
        op([=](cudaStream_t st) {
            group_norm_bwd(xp, C, nullptr, 0, dy, ab, mr, Bn, HW, 32, 1, ws, dx, nullptr, nullptr, st);
            gn_bwd_film_params(ws, Bn, HW, C, film_p, tp, gamma, beta, d_film_p, *gg, *gb, st);
            return (int)cudaGetLastError();
        }, 3);
    }

    void resblock_bwd(const Rec& r, float* d_film) {
        cur_label = "bwd " + r.p;
        const Act &xa = r.xa, &xb = r.xb;
        const int H = xa.H, W = xa.W, Ho = r.Ho, Wo = r.Wo, Cin = xa.C + xb.C, Cout = r.Cout, Bn = B;
        const long long rows_in = (long long)B * H * W, rows = (long long)B * Ho * Wo;
        const bool resample = r.mode == L_DOWN || r.mode == L_UP;
        const bool skip_conv = Cin != Cout;
        const bf16* xs = resample ? r.xp : xa.p;   // skip-path input at the output resolution
        const bf16* conv_src = resample ? r.gp : r.g1;
        const bf16* dO = complete_grad(r.out);
        bias_grad(dO, rows, Cout, r.p + ".out_layers.3.bias", skip_conv ? r.p + ".skip_connection.bias" : "");
        wgrad_any(dO, Cout, r.g2, Cout, Ho, Wo, Cout, Cout, 9, r.p + ".out_layers.3.weight", Cout, 0);
        if (skip_conv) {
            wgrad_any(dO, Cout, xs, xa.C, Ho, Wo, Cout, xa.C, 1, r.p + ".skip_connection.weight", Cin, 0);
            if (xb.C) wgrad_any(dO, Cout, xb.p, xb.C, Ho, Wo, Cout, xb.C, 1, r.p + ".skip_connection.weight", Cin, xa.C);
        }
        bf16* dG2 = (bf16*)scratch(0, (size_t)rows * Cout * 2);
        dgrad(r.p + ".out_layers.3", {r.p + ".out_layers.3.weight"}, {dO}, {Cout}, {9}, Ho, Wo, Cout, dG2, nullptr);
        {
            Plan* pl = &plan;
            const long long n = rows * Cout;
            const unsigned sid = r.drop_stream;
            op([=](cudaStream_t st) {
                if (pl->dropout_p > 0.f) dropout_bf16(dG2, n, pl->dropout_p, pl->dropout_seed, sid, nullptr, st);
                return (int)cudaGetLastError();
            });
        }
        bf16* dH1 = (bf16*)scratch(1, (size_t)rows * Cout * 2);
        if (film_mode) {
            gn_bwd_film(r.p + ".out_layers.0", Act{r.h1, Cout, Ho, Wo}, dG2, r.n2, dH1, film + r.film_off, d_film + r.film_off);
        } else {
            gn_bwd(r.p + ".out_layers.0", Act{r.h1, Cout, Ho, Wo}, Act{}, dG2, r.n2, 1, dH1);
            float* dst = d_film + r.film_off;
            const int ld = TP, HW = Ho * Wo;
            op([=](cudaStream_t st) {
                colsum_per_image(dH1, Bn, HW, Cout, dst, ld, st);
                return (int)cudaGetLastError();
            });
        }
        bias_grad(dH1, rows, Cout, r.p + ".in_layers.2.bias");
        wgrad_any(dH1, Cout, conv_src, Cin, Ho, Wo, Cout, Cin, 9, r.p + ".in_layers.2.weight", Cin, 0);
        bf16* dGp = (bf16*)scratch(2, (size_t)rows * Cin * 2);
        dgrad(r.p + ".in_layers.2", {r.p + ".in_layers.2.weight"}, {dH1}, {Cout}, {9}, Ho, Wo, Cin, dGp, nullptr);
        const bf16* dG1 = dGp;
        if (resample) {
            bf16* t = (bf16*)scratch(3, (size_t)rows_in * Cin * 2);
            const int mode = r.mode;
            op([=](cudaStream_t st) {
                if (mode == L_DOWN) avgpool2_bwd(dGp, t, Bn, H, W, Cin, st);
                else sumpool2(dGp, t, Bn, H, W, Cin, st);
                return (int)cudaGetLastError();
            });
            dG1 = t;
        }
        bf16* dX = (bf16*)scratch(7, (size_t)rows_in * Cin * 2);
        gn_bwd(r.p + ".in_layers.0", xa, xb, dG1, r.n1, 1, dX);
        if (!resample) {
            if (skip_conv) {
                // + skip_connection^T(dO), accumulated in place (every element is read, then written, by the same thread)
                dgrad(r.p + ".skip_connection", {r.p + ".skip_connection.weight"}, {dO}, {Cout}, {1}, H, W, Cin, dX, dX);
            } else {
                op([=](cudaStream_t st) {
                    accum_bf16(dX, dO, Cout, rows, Cout, 0, st);
                    return (int)cudaGetLastError();
                });
            }
        } else {
            // skip path through the resampled input: d xs (output resolution) -> pool / un-pool backward -> + dX
            const bf16* dXs = dO;
            if (skip_conv) {
                bf16* t = (bf16*)scratch(2, (size_t)rows * Cin * 2);
                dgrad(r.p + ".skip_connection", {r.p + ".skip_connection.weight"}, {dO}, {Cout}, {1}, Ho, Wo, Cin, t, nullptr);
                dXs = t;
            }
            bf16* t2 = (bf16*)scratch(3, (size_t)rows_in * Cin * 2);
            const int mode = r.mode;
            op([=](cudaStream_t st) {
                if (mode == L_DOWN) avgpool2_bwd(dXs, t2, Bn, H, W, Cin, st);
                else sumpool2(dXs, t2, Bn, H, W, Cin, st);
                accum_bf16(dX, t2, Cin, rows_in, Cin, 0, st);
                return (int)cudaGetLastError();
            }, 2);
        }
        accumulate(xa, dX, Cin);
        if (xb.C) accumulate(xb, dX + xa.C, Cin);
    }

    void attn_bwd(const Rec& r) {
        cur_label = "bwd " + r.p;
        const Act& x = r.xa;
        const int C = x.C, H = x.H, W = x.W, S = H * W, Bn = B, heads = r.heads, dh = C / heads;
        const long long rows = (long long)B * S;
        const float scale = 1.f / sqrtf((float)dh);
        const bf16* dO = complete_grad(r.out);
        accumulate(x, dO, C);  // residual path
        bias_grad(dO, rows, C, r.p + ".proj_out.bias");
        wgrad_any(dO, C, r.o, C, H, W, C, C, 1, r.p + ".proj_out.weight", C, 0);
        bf16* d_o = (bf16*)scratch(0, (size_t)rows * C * 2);
        dgrad(r.p + ".proj_out", {r.p + ".proj_out.weight"}, {dO}, {C}, {1}, H, W, C, d_o, nullptr);
        bf16* dqkv = (bf16*)scratch(1, (size_t)rows * 3 * C * 2);
        bf16* qkv = r.qkv;
        if (S <= 64) {
            op([=](cudaStream_t st) {
                attn_small_bwd_heads(qkv, d_o, dqkv, Bn, heads, S, dh, scale, st);
                return (int)cudaGetLastError();
            });
        } else {
            float* sc = (float*)scratch(2, (size_t)B * S * S * sizeof(float));  // scores, then dP
            bf16* P = (bf16*)scratch(8, (size_t)B * S * S * 2);
            bf16* dS = (bf16*)scratch(3, (size_t)B * S * S * 2);
            bf16* T1 = (bf16*)scratch(9, (size_t)B * S * S * 2);    // P^T / dS^T
            bf16* T2 = (bf16*)scratch(10, (size_t)B * S * dh * 2);  // dO_h^T / K_h^T / Q_h^T  [B][dh][S]
            const long long qs = (long long)S * 3 * C;
            for (int h = 0; h < heads; ++h) {
                const bf16 *q = qkv + h * dh, *k = qkv + C + h * dh, *v = qkv + 2 * C + h * dh;
                const bf16* doh = d_o + h * dh;
                bf16 *dq = dqkv + h * dh, *dk = dqkv + C + h * dh, *dv = dqkv + 2 * C + h * dh;
                // P = softmax(scale q k^T)  (recomputed: the forward's flash kernel never materialises it)
                bgemm(q, dh, 3 * C, S, k, S, 3 * C, qs, sc, S, (long long)S * S, true, scale, false);
                op([=](cudaStream_t st) {
                    softmax_rows(sc, P, (long long)Bn * S, S, st);
                    return (int)cudaGetLastError();
                });
                // dP = dO v^T ; dS = P (dP - rowsum(P dP)) scale
                bgemm(doh, dh, C, S, v, S, 3 * C, qs, sc, S, (long long)S * S, true, 1.f, false);
                op([=](cudaStream_t st) {
                    softmax_bwd_rows(P, sc, dS, (long long)Bn * S, S, scale, st);
                    transpose_bf16_batched(P, S, (long long)S * S, T1, S, S, Bn, st);
                    transpose_bf16_batched(doh, C, (long long)S * C, T2, S, dh, Bn, st);
                    return (int)cudaGetLastError();
                }, 3);
                // dV = P^T dO
                bgemm(T1, S, S, S, T2, dh, S, (long long)dh * S, dv, 3 * C, qs, false, 1.f, false);
                // dQ = dS K
                op([=](cudaStream_t st) {
                    transpose_bf16_batched(k, 3 * C, qs, T2, S, dh, Bn, st);
                    return (int)cudaGetLastError();
                });
                bgemm(dS, S, S, S, T2, dh, S, (long long)dh * S, dq, 3 * C, qs, false, 1.f, false);
                // dK = dS^T Q
                op([=](cudaStream_t st) {
                    transpose_bf16_batched(dS, S, (long long)S * S, T1, S, S, Bn, st);
                    transpose_bf16_batched(q, 3 * C, qs, T2, S, dh, Bn, st);
                    return (int)cudaGetLastError();
                }, 2);
                bgemm(T1, S, S, S, T2, dh, S, (long long)dh * S, dk, 3 * C, qs, false, 1.f, false);
            }
        }
        bias_grad(dqkv, rows, 3 * C, r.p + ".qkv.bias");
        wgrad_any(dqkv, 3 * C, r.hn, C, H, W, 3 * C, C, 1, r.p + ".qkv.weight", C, 0);
        bf16* d_hn = (bf16*)scratch(0, (size_t)rows * C * 2);
        dgrad(r.p + ".qkv", {r.p + ".qkv.weight"}, {dqkv}, {3 * C}, {1}, H, W, C, d_hn, nullptr);
        bf16* dX = (bf16*)scratch(2, (size_t)rows * C * 2);
        gn_bwd(r.p + ".norm", x, Act{}, d_hn, r.n1, 0, dX);
        accumulate(x, dX, C);
    }

    // ---------------------------------------------------------------- whole network
    struct ResInfo {
        std::string p;
        int cout;
    };

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int mc = a.ch, R = a.resolution;
        ted = 4 * mc;
        gn_eps = 1e-5f;  // models/cm/nn.py:109-116 (GroupNorm32 default eps)
        film_mode = a.use_scale_shift_norm != 0;
        Plan* pl = &plan;
        const int Bn = B;
        if (!a.resblock_updown) {
            fail("ADM U-Net: only resblock_updown=True is built (every DxMI EDM config uses it)");
            return;
        }
        if (a.in_channels != 3 || a.out_channels != 3 || (R * R) % 128 || B > 1024)
            fail("ADM training plan: needs 3 -> 3 channels, H*W % 128 == 0, batch <= 1024");
        plan.eps = (float*)alloc((size_t)B * a.out_channels * R * R * sizeof(float));
        plan.tbuf = (float*)alloc((size_t)B * sizeof(float));
        plan.coef = (float*)alloc((size_t)B * 8 * sizeof(float));

        BlockList inputs, outputs;
        int mid = 0;
        adm_layout(a, inputs, outputs, &mid);
        const std::vector<Layer> middle{{L_RES, mid, mid}, {L_ATTN, mid, mid}, {L_RES, mid, mid}};

        // ================================================================= forward
        emit_bwd = false;
        std::vector<ResInfo> rbs;
        std::vector<int> film_offs;
        TP = 0;
        auto collect = [&](const std::string& prefix, const std::vector<Layer>& layers) {
            for (size_t j = 0; j < layers.size(); ++j)
                if (layers[j].kind == L_RES || layers[j].kind == L_DOWN || layers[j].kind == L_UP) {
                    rbs.push_back({prefix + "." + std::to_string(j), layers[j].cout});
                    film_offs.push_back(TP);
                    TP += (film_mode ? 2 : 1) * layers[j].cout;
                }
        };
        for (size_t i = 1; i < inputs.size(); ++i) collect("input_blocks." + std::to_string(i), inputs[i]);
        collect("middle_block", middle);
        for (size_t i = 0; i < outputs.size(); ++i) collect("output_blocks." + std::to_string(i), outputs[i]);

        float* te = (float*)alloc((size_t)B * mc * 4);
        float* t1 = (float*)alloc((size_t)B * ted * 4);
        float* emb = (float*)alloc((size_t)B * ted * 4);
        film = (float*)alloc((size_t)B * TP * 4);
        const float* w0 = f32("time_embed.0.weight");
        const float* b0 = f32("time_embed.0.bias");
        const float* w2 = f32("time_embed.2.weight");
        const float* b2 = f32("time_embed.2.bias");
        const float* table = a.num_classes > 0 ? f32("label_emb.weight") : nullptr;
        const bool has_labels = a.num_classes > 0;
        {
            const int tc = ted;
            cur_label = "emb";
            op([=](cudaStream_t st) {
                // models/cm/unet.py:775-779: emb = time_embed(timestep_embedding(t)) (+ label_emb(y)), all fp32
                timestep_embedding(pl->t, te, Bn, mc, 1, st);
                linear_f32(te, mc, w0, b0, t1, tc, Bn, mc, tc, 0, 0, st);
                linear_f32(t1, tc, w2, b2, emb, tc, Bn, tc, tc, 2, 0, st);
                if (has_labels) {
                    if (!pl->y) return (int)cudaErrorInvalidValue;  // class-conditional net needs labels
                    embedding_add(emb, table, (const long long*)pl->y, Bn, tc, st);
                }
                return (int)cudaGetLastError();
            }, has_labels ? 4 : 3);
            std::vector<std::string> wk, bk;
            for (auto& rb : rbs) {
                wk.push_back(rb.p + ".emb_layers.1.weight");
                bk.push_back(rb.p + ".emb_layers.1.bias");
            }
            batched_emb_projection(emb, ted, "emb_layers", wk, bk, film, TP);
        }
        const int ch0 = inputs[0][0].cout;
        Act h = mk(ch0, R, R);
        if (ch0 % 32 || ch0 > 256 || h.stats_P < R * R / 128 || h.stats_halo) fail("ADM input conv: unsupported geometry");
        h.stats_P = R * R / 128;  // conv3x3_first_k publishes one partial per 128-pixel tile
        {
            const float* w = f32("input_blocks.0.0.weight");
            const float* b = f32("input_blocks.0.0.bias");
            bf16* o = h.p;
            float* hst = h.stats;
            cur_label = "input conv";
            op([=](cudaStream_t st) {
                conv3x3_first(pl->x, nullptr, w, b, o, hst, Bn, 3, R, R, ch0, 0, st);
                return (int)cudaGetLastError();
            });
            Rec r;
            r.kind = 0;
            r.out = h;
            tape.push_back(r);
        }
        int ri = 0;
        auto run_layers = [&](const std::string& prefix, const std::vector<Layer>& layers, Act hh, Act skip) {
            for (size_t j = 0; j < layers.size(); ++j) {
                const Layer& L = layers[j];
                const std::string p = prefix + "." + std::to_string(j);
                if (L.kind == L_ATTN) {
                    hh = attention(p, hh);
                } else {
                    hh = resblock(p, hh, j == 0 ? skip : Act{}, L.cout, L.kind, film_offs[ri]);
                    ++ri;
                }
            }
            return hh;
        };
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
        Rec head;
        head.kind = 5;
        head.xa = h;
        head.g1 = act_alloc(h.C, R, R);
        head.n1 = gn_fwd(h, Act{}, "out.0", 1, head.g1);
        cur_label = "out conv";
        conv_out_nchw(head.g1, h.C, R, R, "out.2.weight", "out.2.bias", a.out_channels);
        tape.push_back(head);

        // ================================================================= backward
        emit_bwd = true;
        float* d_film = (float*)alloc((size_t)B * TP * sizeof(float));
        for (int i = (int)tape.size() - 1; i >= 0; --i) {
            const Rec& r = tape[i];
            if (r.kind == 5) {
                cur_label = "bwd head";
                const int C = r.xa.C;
                if (C > 256) fail("ADM training: the output convolution's input width must be <= 256");
                float* w_t = nullptr;
                if (!dry) {
                    bool fresh = false;
                    w_t = (float*)derived_buf("wT:out.2.first", (size_t)C * 27 * sizeof(float), &fresh);
                    if (fresh) {
                        Net* np = &net;
                        net.pack_jobs.push_back([np, w_t, C](cudaStream_t st) {
                            const Bound& bb = np->bound["out.2.weight"];
                            conv_out_transpose_weights((const float*)bb.ptr, w_t, C, st);
                            count_launches(1);
                        });
                        const Bound* bw = get("out.2.weight");
                        if (bw && bw->dtype != DXMI_F32) fail("out.2.weight must be fp32 for training (models/cm/unet.py:738-742 keeps the head in fp32)");
                    }
                }
                bf16* dG = (bf16*)scratch(0, (size_t)B * R * R * C * 2);
                float* wsf = (float*)scratch(4, (size_t)B * (R / 4) * C * 27 * sizeof(float));
                float* T = (float*)alloc((size_t)C * 27 * sizeof(float));
                float* wsb = (float*)alloc((size_t)B * 3 * sizeof(float));
                float **gw = gslot("out.2.weight"), **gb = gslot("out.2.bias");
                const bf16* g = r.g1;
                op([=](cudaStream_t st) {
                    if (*gb) sum_nchw_channels(pl->dout, Bn, 3, R * R, wsb, *gb, st);
                    if (*gw) {
                        conv_first_wgrad(g, pl->dout, wsf, T, Bn, R, R, C, st);
                        conv_out_wgrad_fix(T, *gw, C, st);
                    }
                    conv3x3_first(pl->dout, nullptr, w_t, nullptr, dG, nullptr, Bn, 3, R, R, C, 0, st);
                    return (int)cudaGetLastError();
                }, 5);
                bf16* dH = (bf16*)scratch(1, (size_t)B * R * R * C * 2);
                gn_bwd("out.0", r.xa, Act{}, dG, r.n1, 1, dH);
                accumulate(r.xa, dH, C);
            } else if (r.kind == 1) {
                resblock_bwd(r, d_film);
            } else if (r.kind == 2) {
                attn_bwd(r);
            } else if (r.kind == 0) {
                cur_label = "bwd input conv";
                const bf16* dZ = complete_grad(r.out);
                bias_grad(dZ, (long long)B * R * R, ch0, "input_blocks.0.0.bias");
                float* wsf = (float*)scratch(4, (size_t)B * (R / 4) * ch0 * 27 * sizeof(float));
                float** g = gslot("input_blocks.0.0.weight");
                op([=](cudaStream_t st) {
                    if (!*g) return 0;
                    conv_first_wgrad(dZ, pl->x, wsf, *g, Bn, R, R, ch0, st);
                    return (int)cudaGetLastError();
                }, 2);
            }
        }
        // ---- embedding path: d_film [B, TP] -> emb_layers stack, time_embed.2 / .0, label_emb
        cur_label = "bwd emb";
        float* d_se = (float*)alloc((size_t)B * ted * sizeof(float));  // grad w.r.t. silu(emb), then emb
        float* d_s1 = (float*)alloc((size_t)B * ted * sizeof(float));  // grad w.r.t. silu(t1), then t1
        float* s_emb = (float*)alloc((size_t)B * ted * sizeof(float));
        float* s_t1 = (float*)alloc((size_t)B * ted * sizeof(float));
        {
            const int tc = ted, tp = TP;
            op([=](cudaStream_t st) {
                silu_f32(emb, s_emb, (long long)Bn * tc, st);
                silu_f32(t1, s_t1, (long long)Bn * tc, st);
                return (int)cudaGetLastError();
            }, 2);
            LinearStack base{};
            std::vector<float**> gws, gbs;
            if ((int)rbs.size() > LinearStack::MAX) fail("too many ResBlocks for the stacked emb_layers backward");
            base.n_layers = (int)rbs.size();
            for (size_t i = 0; i < rbs.size() && i < (size_t)LinearStack::MAX; ++i) {
                base.off[i] = film_offs[i];
                base.W[i] = f32(rbs[i].p + ".emb_layers.1.weight");
                gws.push_back(gslot(rbs[i].p + ".emb_layers.1.weight"));
                gbs.push_back(gslot(rbs[i].p + ".emb_layers.1.bias"));
            }
            base.off[base.n_layers] = tp;
            op([=](cudaStream_t st) {
                LinearStack ls = base;
                for (int i = 0; i < ls.n_layers; ++i) {
                    ls.dW[i] = *gws[i];
                    ls.db[i] = *gbs[i];
                }
                linear_stack_bwd_w(d_film, tp, s_emb, tc, ls, Bn, tc, st);
                linear_stack_bwd_x(d_film, tp, ls, d_se, tc, Bn, tc, st);
                return (int)cudaGetLastError();
            }, 2);
            float **g2w = gslot("time_embed.2.weight"), **g2b = gslot("time_embed.2.bias");
            float **g0w = gslot("time_embed.0.weight"), **g0b = gslot("time_embed.0.bias");
            float** gl = has_labels ? gslot("label_emb.weight") : nullptr;
            const int ncls = a.num_classes;
            op([=](cudaStream_t st) {
                silu_bwd_mul(d_se, emb, (long long)Bn * tc, st);  // -> grad w.r.t. emb
                if (gl && *gl) embedding_grad(d_se, (const long long*)pl->y, *gl, Bn, tc, ncls, st);
                linear_bwd_w(d_se, tc, s_t1, tc, 0, *g2w, *g2b, Bn, tc, tc, st);
                linear_bwd_x(d_se, tc, w2, d_s1, tc, Bn, tc, tc, 0, st);
                silu_bwd_mul(d_s1, t1, (long long)Bn * tc, st);  // -> grad w.r.t. t1
                linear_bwd_w(d_s1, tc, te, mc, 0, *g0w, *g0b, Bn, tc, mc, st);
                return (int)cudaGetLastError();
            }, 7);
        }
        emit_bwd = false;
    }
};

int build_adm_train_plan(Net& net, Plan& plan) {
    if (net.a.precision != 0) {
        engine_set_error("training plans run in bf16 mode only");
        return -27;
    }
    return build_two_pass<AdmTrainBuilder>(net, plan);
}

}  // namespace dxmi
