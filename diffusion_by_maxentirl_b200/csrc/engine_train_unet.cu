// Training plan of the DDPM U-Net (models/DxMI/unet_small.py:194-332 under autograd; SURVEY 8a row a9, trainer.py:348-389
// update_sampler): a forward pass that keeps what the backward needs, and the backward pass from d loss / d eps to every
// parameter gradient, as two static launch lists.  Dropout must be 0 ("dropout forced to 0 for parity").
//
// Reverse-mode structure: the forward records a tape (conv_in, ResnetBlocks, AttnBlocks, Downsample, Upsample, head); the
// backward walks it in reverse.  Every activation that other ops consume owns a bf16 gradient buffer; contributions (next op,
// skip-connection concat, residual) are accumulated with accum_bf16 - all consumers of an activation come later in the forward,
// so its gradient is complete when the reverse walk reaches its producer.
//   dense parts : data gradients = forward implicit-GEMM kernels on transposed, tap-flipped weights; weight gradients = the
//                 MN-major split-K tcgen05 kernel (wgrad_tc.cu), per channel slice for concatenated inputs
//   GroupNorm   : gn_bwd_* (kernels_bwd.cu) from the forward's saved per-(image, channel) affine and (mean, rstd)
//   attention   : P = softmax(q k^T) is kept; dP = dO V^T, dV = P^T dO, dQ = dS K, dK = dS^T Q as batched tensor-core GEMMs on
//                 explicitly transposed operands (seq 256), or one SIMT CTA per image (seq <= 64)
//   Downsample  : stride-2 pad-(0,1,0,1) conv: dgrad and wgrad are the standard pad-1 operators on dY zero-inserted at odd positions
//   conv_out    : dgrad = the first-conv kernel on transposed weights, wgrad = the first-conv wgrad with taps flipped
//   temb path   : per-image column sums -> fp32 Linear backward kernels
#include "train_builder.cuh"

namespace dxmi {

namespace {
bool has_attn_t(const dxmi_arch_desc& a, int v) {
    for (int i = 0; i < a.n_attn; ++i)
        if (a.attn_resolutions[i] == v) return true;
    return false;
}
}  // namespace

struct DdpmTrainBuilder : TrainBuilder {
    using TrainBuilder::TrainBuilder;

    struct Rec {
        int kind = 0;  // 0 conv_in, 1 resblock, 2 attn, 3 down, 4 up, 5 head
        std::string p;
        Act xa, xb, out;
        int Cout = 0;
        bf16 *g1 = nullptr, *h1 = nullptr, *g2 = nullptr;
        GnSave n1, n2;
        int tp_off = 0;
        unsigned drop_stream = 0;
        // attention
        bf16 *hn = nullptr, *qkv = nullptr, *P = nullptr, *o = nullptr;
        // up
        bf16* up = nullptr;
    };
    std::vector<Rec> tape;
    float *tproj = nullptr, *temb = nullptr, *t1 = nullptr, *te = nullptr;
    int TP = 0, temb_ch = 0;


    Act resblock(const std::string& p, Act xa, Act xb, int Cout, int tp_off) {
        cur_label = p;
        Rec r;
        r.kind = 1;
        r.p = p;
        r.xa = xa;
        r.xb = xb;
        r.Cout = Cout;
        r.tp_off = tp_off;
        const int H = xa.H, W = xa.W, Cin = xa.C + xb.C;
        r.g1 = act_alloc(Cin, H, W);
        r.n1 = gn_fwd(xa, xb, p + ".norm1", 1, r.g1);
        Act h1 = conv3(p + ".conv1", r.g1, Cin, H, W, Cout, tproj + tp_off, TP, nullptr);
        r.h1 = h1.p;
        r.g2 = act_alloc(Cout, H, W);
        r.n2 = gn_fwd(h1, Act{}, p + ".norm2", 1, r.g2);
        {
            // self.dropout(h) between swish(norm2) and conv2 (unet_small.py:126-127): in place, mask regenerated in the backward
            Plan* pl = &plan;
            bf16* g2 = r.g2;
            const long long n = (long long)B * H * W * Cout;
            const unsigned sid = (unsigned)tape.size();
            r.drop_stream = sid;
            op([=](cudaStream_t st) {
                if (pl->dropout_p > 0.f) dropout_bf16(g2, n, pl->dropout_p, pl->dropout_seed, sid, nullptr, st);
                return (int)cudaGetLastError();
            });
        }
        Act out = mk(Cout, H, W);
        {
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, r.g2, Cout, Cout);
            add_seg(d, 0, 9);
            long long K = 9LL * Cout;
            if (Cin != Cout) {
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
        r.out = out;
        tape.push_back(r);
        return out;
    }

    Act attn(const std::string& p, Act x) {
        cur_label = p;
        Rec r;
        r.kind = 2;
        r.p = p;
        r.xa = x;
        const int C = x.C, H = x.H, W = x.W, S = H * W;
        const float scale = 1.f / sqrtf((float)C);
        r.hn = act_alloc(C, H, W);
        r.n1 = gn_fwd(x, Act{}, p + ".norm", 0, r.hn);
        r.qkv = (bf16*)alloc((size_t)B * S * 3 * C * sizeof(bf16));
        {
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, r.hn, C, C);
            add_seg(d, 0, 1);
            d.b_ptr = packed_rows(p + ".qkv", {{{p + ".q.weight", 0, C}}, {{p + ".k.weight", 0, C}}, {{p + ".v.weight", 0, C}}}, nullptr,
                                  nullptr);
            d.b_rows = 3 * C;
            d.b_ld = C;
            d.bias = concat_f32(p + ".qkv.bias", {p + ".q.bias", p + ".k.bias", p + ".v.bias"});
            d.out = r.qkv;
            d.ldo = 3 * C;
            gemm(d);
        }
        r.o = act_alloc(C, H, W);
        const int Bn = B;
        bf16 *qkv = r.qkv, *o = r.o;
        if (S <= 64) {
            op([=](cudaStream_t st) {
                attn_small(qkv, qkv + C, qkv + 2 * C, 3 * C, o, C, Bn, 1, S, C, scale, st);
                return (int)cudaGetLastError();
            });
        } else if (S == 256 || S == 128) {
            r.P = (bf16*)alloc((size_t)B * S * S * sizeof(bf16));
            bf16* vT = (bf16*)scratch(1, (size_t)B * C * S * sizeof(bf16));
            op([=](cudaStream_t st) {
                transpose_bf16_batched(qkv + 2 * C, 3 * C, (long long)S * 3 * C, vT, S, C, Bn, st);
                return (int)cudaGetLastError();
            });
            bgemm(qkv, C, 3 * C, S, qkv + C, S, 3 * C, (long long)S * 3 * C, r.P, S, (long long)S * S, false, scale, true);
            bgemm(r.P, S, S, S, vT, C, S, (long long)C * S, o, C, (long long)S * C, false, 1.f, false);
        } else {
            fail("DDPM training attention: unsupported sequence length");
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


    void resblock_bwd(const Rec& r, float* d_tproj) {
        cur_label = "bwd " + r.p;
        const Act &xa = r.xa, &xb = r.xb;
        const int H = xa.H, W = xa.W, Cin = xa.C + xb.C, Cout = r.Cout, Bn = B;
        const long long rows = (long long)B * H * W;
        const bf16* dO = complete_grad(r.out);
        const bool nin = Cin != Cout;
        bias_grad(dO, rows, Cout, r.p + ".conv2.bias", nin ? r.p + ".nin_shortcut.bias" : "");
        wgrad_sliced(dO, Cout, r.g2, Cout, H, W, Cout, Cout, 9, r.p + ".conv2.weight", Cout, 0);
        if (nin) {
            wgrad_sliced(dO, Cout, xa.p, xa.C, H, W, Cout, xa.C, 1, r.p + ".nin_shortcut.weight", Cin, 0);
            if (xb.C) wgrad_sliced(dO, Cout, xb.p, xb.C, H, W, Cout, xb.C, 1, r.p + ".nin_shortcut.weight", Cin, xa.C);
        }
        bf16* dG2 = (bf16*)scratch(0, (size_t)rows * Cout * 2);
        dgrad(r.p + ".conv2", {r.p + ".conv2.weight"}, {dO}, {Cout}, {9}, H, W, Cout, dG2, nullptr);
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
        gn_bwd(r.p + ".norm2", Act{r.h1, Cout, H, W}, Act{}, dG2, r.n2, 1, dH1);
        {
            float* dst = d_tproj + r.tp_off;
            const int ld = TP, HW = H * W;
            op([=](cudaStream_t st) {
                colsum_per_image(dH1, Bn, HW, Cout, dst, ld, st);
                return (int)cudaGetLastError();
            });
        }
        bias_grad(dH1, rows, Cout, r.p + ".conv1.bias");
        wgrad_sliced(dH1, Cout, r.g1, Cin, H, W, Cout, Cin, 9, r.p + ".conv1.weight", Cin, 0);
        bf16* dG1 = (bf16*)scratch(2, (size_t)rows * Cin * 2);
        dgrad(r.p + ".conv1", {r.p + ".conv1.weight"}, {dH1}, {Cout}, {9}, H, W, Cin, dG1, nullptr);
        bf16* dX = (bf16*)scratch(3, (size_t)rows * Cin * 2);
        gn_bwd(r.p + ".norm1", xa, xb, dG1, r.n1, 1, dX);
        if (nin) {
            // + nin_shortcut^T(dO), accumulated in place (every element is read, then written, by the same thread)
            dgrad(r.p + ".nin_shortcut", {r.p + ".nin_shortcut.weight"}, {dO}, {Cout}, {1}, H, W, Cin, dX, dX);
        } else {
            op([=](cudaStream_t st) {
                accum_bf16(dX, dO, Cout, rows, Cout, 0, st);
                return (int)cudaGetLastError();
            });
        }
        accumulate(xa, dX, Cin);
        if (xb.C) accumulate(xb, dX + xa.C, Cin);
    }

    void attn_bwd(const Rec& r) {
        cur_label = "bwd " + r.p;
        const Act& x = r.xa;
        const int C = x.C, H = x.H, W = x.W, S = H * W, Bn = B;
        const long long rows = (long long)B * S;
        const float scale = 1.f / sqrtf((float)C);
        const bf16* dO = complete_grad(r.out);
        accumulate(x, dO, C);  // residual path
        bias_grad(dO, rows, C, r.p + ".proj_out.bias");
        wgrad_sliced(dO, C, r.o, C, H, W, C, C, 1, r.p + ".proj_out.weight", C, 0);
        bf16* d_o = (bf16*)scratch(0, (size_t)rows * C * 2);
        dgrad(r.p + ".proj_out", {r.p + ".proj_out.weight"}, {dO}, {C}, {1}, H, W, C, d_o, nullptr);
        bf16* dqkv = (bf16*)scratch(1, (size_t)rows * 3 * C * 2);
        bf16* qkv = r.qkv;
        if (S <= 64) {
            op([=](cudaStream_t st) {
                attn_small_bwd(qkv, d_o, dqkv, Bn, S, C, scale, st);
                return (int)cudaGetLastError();
            });
        } else {
            float* dP = (float*)scratch(2, (size_t)B * S * S * sizeof(float));
            bf16* dS = (bf16*)scratch(3, (size_t)B * S * S * 2);
            bf16* T1 = (bf16*)scratch(6, (size_t)B * S * (S > C ? S : C) * 2);  // transposed operand #1
            bf16* T2 = (bf16*)scratch(7, (size_t)B * S * (S > C ? S : C) * 2);  // transposed operand #2
            bf16* P = r.P;
            // dP = dO . V^T
            bgemm(d_o, C, C, S, qkv + 2 * C, S, 3 * C, (long long)S * 3 * C, dP, S, (long long)S * S, true, 1.f, false);
            op([=](cudaStream_t st) {
                softmax_bwd_rows(P, dP, dS, (long long)Bn * S, S, scale, st);
                // dV = P^T . dO : A = P^T, B = dO^T
                transpose_bf16_batched(P, S, (long long)S * S, T1, S, S, Bn, st);
                transpose_bf16_batched(d_o, C, (long long)S * C, T2, S, C, Bn, st);
                return (int)cudaGetLastError();
            }, 3);
            bgemm(T1, S, S, S, T2, C, S, (long long)C * S, dqkv + 2 * C, 3 * C, (long long)S * 3 * C, false, 1.f, false);
            // dQ = dS . K : B = K^T
            op([=](cudaStream_t st) {
                transpose_bf16_batched(qkv + C, 3 * C, (long long)S * 3 * C, T2, S, C, Bn, st);
                return (int)cudaGetLastError();
            });
            bgemm(dS, S, S, S, T2, C, S, (long long)C * S, dqkv, 3 * C, (long long)S * 3 * C, false, 1.f, false);
            // dK = dS^T . Q : A = dS^T, B = Q^T
            op([=](cudaStream_t st) {
                transpose_bf16_batched(dS, S, (long long)S * S, T1, S, S, Bn, st);
                transpose_bf16_batched(qkv, 3 * C, (long long)S * 3 * C, T2, S, C, Bn, st);
                return (int)cudaGetLastError();
            }, 2);
            bgemm(T1, S, S, S, T2, C, S, (long long)C * S, dqkv + C, 3 * C, (long long)S * 3 * C, false, 1.f, false);
        }
        // q / k / v projections
        {
            float* ws = (float*)scratch(5, (size_t)colsum_ws_floats(rows, 3 * C) * sizeof(float));
            float* tmp = (float*)alloc((size_t)3 * C * sizeof(float));
            float **gq = gslot(r.p + ".q.bias"), **gk = gslot(r.p + ".k.bias"), **gv = gslot(r.p + ".v.bias");
            op([=](cudaStream_t st) {
                colsum_bf16(dqkv, rows, 3 * C, ws, tmp, st);
                if (*gq) cudaMemcpyAsync(*gq, tmp, (size_t)C * 4, cudaMemcpyDeviceToDevice, st);
                if (*gk) cudaMemcpyAsync(*gk, tmp + C, (size_t)C * 4, cudaMemcpyDeviceToDevice, st);
                if (*gv) cudaMemcpyAsync(*gv, tmp + 2 * C, (size_t)C * 4, cudaMemcpyDeviceToDevice, st);
                return (int)cudaGetLastError();
            }, 2);
        }
        wgrad_sliced(dqkv, 3 * C, r.hn, C, H, W, C, C, 1, r.p + ".q.weight", C, 0);
        wgrad_sliced(dqkv + C, 3 * C, r.hn, C, H, W, C, C, 1, r.p + ".k.weight", C, 0);
        wgrad_sliced(dqkv + 2 * C, 3 * C, r.hn, C, H, W, C, C, 1, r.p + ".v.weight", C, 0);
        bf16* d_hn = (bf16*)scratch(0, (size_t)rows * C * 2);
        dgrad(r.p + ".qkv", {r.p + ".q.weight", r.p + ".k.weight", r.p + ".v.weight"}, {dqkv}, {3 * C}, {1}, H, W, C, d_hn, nullptr);
        bf16* dX = (bf16*)scratch(2, (size_t)rows * C * 2);
        gn_bwd(r.p + ".norm", x, Act{}, d_hn, r.n1, 0, dX);
        accumulate(x, dX, C);
    }

    void build() {
        const dxmi_arch_desc& a = net.a;
        const int ch = a.ch, R = a.resolution;
        temb_ch = 4 * ch;
        Plan* pl = &plan;
        const int Bn = B;
        if (a.in_channels != 3 || a.out_channels != 3 || ch % 128 || ch > 256 || (R * R) % 128 || B > 1024)
            fail("DDPM training plan: needs 3 -> 3 channels, ch in {128, 256}, H*W % 128 == 0, batch <= 1024");
        plan.eps = (float*)alloc((size_t)B * a.out_channels * R * R * sizeof(float));
        plan.tbuf = (float*)alloc((size_t)B * sizeof(float));
        plan.coef = (float*)alloc((size_t)B * 8 * sizeof(float));

        // ================================================================= forward
        emit_bwd = false;
        std::vector<std::string> rb;
        for (int l = 0; l < a.n_levels; ++l)
            for (int b = 0; b < a.num_res_blocks; ++b) rb.push_back("down." + std::to_string(l) + ".block." + std::to_string(b));
        rb.push_back("mid.block_1");
        rb.push_back("mid.block_2");
        for (int l = a.n_levels - 1; l >= 0; --l)
            for (int b = 0; b <= a.num_res_blocks; ++b) rb.push_back("up." + std::to_string(l) + ".block." + std::to_string(b));
        std::vector<int> rb_cout;
        for (int l = 0; l < a.n_levels; ++l)
            for (int b = 0; b < a.num_res_blocks; ++b) rb_cout.push_back(ch * a.ch_mult[l]);
        rb_cout.push_back(ch * a.ch_mult[a.n_levels - 1]);
        rb_cout.push_back(ch * a.ch_mult[a.n_levels - 1]);
        for (int l = a.n_levels - 1; l >= 0; --l)
            for (int b = 0; b <= a.num_res_blocks; ++b) rb_cout.push_back(ch * a.ch_mult[l]);
        TP = 0;
        std::vector<int> tp_offs;
        for (int c : rb_cout) {
            tp_offs.push_back(TP);
            TP += c;
        }
        te = (float*)alloc((size_t)B * ch * 4);
        t1 = (float*)alloc((size_t)B * temb_ch * 4);
        temb = (float*)alloc((size_t)B * temb_ch * 4);
        tproj = (float*)alloc((size_t)B * TP * 4);
        const float* w0 = f32("temb.dense.0.weight");
        const float* b0 = f32("temb.dense.0.bias");
        const float* w1 = f32("temb.dense.1.weight");
        const float* b1 = f32("temb.dense.1.bias");
        {
            std::vector<std::string> wk, bk;
            for (auto& p : rb) {
                wk.push_back(p + ".temb_proj.weight");
                bk.push_back(p + ".temb_proj.bias");
            }
            float *te_ = te, *t1_ = t1, *temb_ = temb;
            const int tc = temb_ch;
            cur_label = "temb";
            op([=](cudaStream_t st) {
                timestep_embedding(pl->t, te_, Bn, ch, 0, st);
                linear_f32(te_, ch, w0, b0, t1_, tc, Bn, ch, tc, 0, 0, st);
                linear_f32(t1_, tc, w1, b1, temb_, tc, Bn, tc, tc, 2, 0, st);
                return (int)cudaGetLastError();
            }, 3);
            batched_emb_projection(temb, temb_ch, "temb_proj", wk, bk, tproj, TP);
        }
        Act h0 = mk(ch, R, R);
        if (h0.stats_P < R * R / 128 || h0.stats_halo) fail("DDPM conv_in: unsupported geometry");
        h0.stats_P = R * R / 128;  // conv3x3_first_k publishes one partial per 128-pixel tile
        {
            const float* w = f32("conv_in.weight");
            const float* b = f32("conv_in.bias");
            bf16* o = h0.p;
            float* hst = h0.stats;
            cur_label = "conv_in";
            op([=](cudaStream_t st) {
                conv3x3_first(pl->x, nullptr, w, b, o, hst, Bn, 3, R, R, ch, 0, st);
                return (int)cudaGetLastError();
            });
            Rec r;
            r.kind = 0;
            r.out = h0;
            tape.push_back(r);
        }
        std::vector<Act> hs{h0};
        int res = R, ri = 0;
        for (int l = 0; l < a.n_levels; ++l) {
            const int cout = ch * a.ch_mult[l];
            const std::string lp = "down." + std::to_string(l);
            for (int b = 0; b < a.num_res_blocks; ++b, ++ri) {
                Act h = resblock(lp + ".block." + std::to_string(b), hs.back(), Act{}, cout, tp_offs[ri]);
                if (has_attn_t(a, res)) h = attn(lp + ".attn." + std::to_string(b), h);
                hs.push_back(h);
            }
            if (l != a.n_levels - 1) {
                Act x = hs.back();
                const std::string p = lp + ".downsample";
                cur_label = p;
                Act out = mk(x.C, x.H / 2, x.W / 2);
                dxmi_gemm_desc d = conv_desc(x.H, x.W);
                d.out_H = x.H / 2;
                d.out_W = x.W / 2;
                d.stride = 2;
                d.rows_per_image = (x.H / 2) * (x.W / 2);
                set_src(d, 0, x.p, x.C, x.C);
                add_seg(d, 0, 9);
                d.b_ptr = packed_rows(p + ".conv", {{{p + ".conv.weight", 0, x.C}}}, nullptr, nullptr);
                d.b_rows = x.C;
                d.b_ld = 9LL * x.C;
                d.bias = f32(p + ".conv.bias");
                d.out = out.p;
                d.ldo = x.C;
                want_stats(d, out);
                gemm(d);
                Rec r;
                r.kind = 3;
                r.p = p;
                r.xa = x;
                r.out = out;
                tape.push_back(r);
                hs.push_back(out);
                res /= 2;
            }
        }
        Act h = hs.back();
        h = resblock("mid.block_1", h, Act{}, h.C, tp_offs[ri++]);
        h = attn("mid.attn_1", h);
        h = resblock("mid.block_2", h, Act{}, h.C, tp_offs[ri++]);
        for (int l = a.n_levels - 1; l >= 0; --l) {
            const int cout = ch * a.ch_mult[l];
            const std::string lp = "up." + std::to_string(l);
            for (int b = 0; b <= a.num_res_blocks; ++b, ++ri) {
                Act skip = hs.back();
                hs.pop_back();
                h = resblock(lp + ".block." + std::to_string(b), h, skip, cout, tp_offs[ri]);
                if (has_attn_t(a, res)) h = attn(lp + ".attn." + std::to_string(b), h);
            }
            if (l != 0) {
                const std::string p = lp + ".upsample";
                cur_label = p;
                const int C = h.C, H2 = h.H * 2, W2 = h.W * 2, H = h.H, W = h.W;
                bf16* up = act_alloc(C, H2, W2);
                const bf16* xp = h.p;
                op([=](cudaStream_t st) {
                    upsample2x(xp, up, Bn, H, W, C, st);
                    return (int)cudaGetLastError();
                });
                Act out = conv3(p + ".conv", up, C, H2, W2, C, nullptr, 0, nullptr);
                Rec r;
                r.kind = 4;
                r.p = p;
                r.xa = h;
                r.up = up;
                r.out = out;
                tape.push_back(r);
                h = out;
                res *= 2;
            }
        }
        Rec head;
        head.kind = 5;
        head.xa = h;
        head.g1 = act_alloc(h.C, R, R);
        head.n1 = gn_fwd(h, Act{}, "norm_out", 1, head.g1);
        cur_label = "conv_out";
        conv_out_nchw(head.g1, h.C, R, R, "conv_out.weight", "conv_out.bias", a.out_channels);
        tape.push_back(head);

        // ================================================================= backward
        emit_bwd = true;
        float* d_tproj = (float*)alloc((size_t)B * TP * sizeof(float));
        for (int i = (int)tape.size() - 1; i >= 0; --i) {
            const Rec& r = tape[i];
            if (r.kind == 5) {
                cur_label = "bwd head";
                const int C = r.xa.C;
                // conv_out: dgrad through the first-conv kernel on transposed weights; wgrad through the first-conv wgrad
                float* w_t = nullptr;
                if (!dry) {
                    bool fresh = false;
                    w_t = (float*)derived_buf("wT:conv_out.first", (size_t)C * 27 * sizeof(float), &fresh);
                    if (fresh) {
                        Net* np = &net;
                        net.pack_jobs.push_back([np, w_t, C](cudaStream_t st) {
                            const Bound& bb = np->bound["conv_out.weight"];
                            conv_out_transpose_weights((const float*)bb.ptr, w_t, C, st);
                            count_launches(1);
                        });
                        const Bound* bw = get("conv_out.weight");
                        if (bw && bw->dtype != DXMI_F32) fail("conv_out.weight must be fp32 for training");
                    }
                }
                bf16* dG = (bf16*)scratch(0, (size_t)B * R * R * C * 2);
                float* wsf = (float*)scratch(4, (size_t)B * (R / 4) * C * 27 * sizeof(float));
                float* T = (float*)alloc((size_t)C * 27 * sizeof(float));
                float* wsb = (float*)alloc((size_t)B * 3 * sizeof(float));
                float **gw = gslot("conv_out.weight"), **gb = gslot("conv_out.bias");
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
                gn_bwd("norm_out", r.xa, Act{}, dG, r.n1, 1, dH);
                accumulate(r.xa, dH, C);
            } else if (r.kind == 1) {
                resblock_bwd(r, d_tproj);
            } else if (r.kind == 2) {
                attn_bwd(r);
            } else if (r.kind == 4) {
                cur_label = "bwd " + r.p;
                const int C = r.xa.C, H2 = r.out.H, W2 = r.out.W, H = r.xa.H, W = r.xa.W;
                const long long rows2 = (long long)B * H2 * W2;
                const bf16* dO = complete_grad(r.out);
                bias_grad(dO, rows2, C, r.p + ".conv.bias");
                wgrad_sliced(dO, C, r.up, C, H2, W2, C, C, 9, r.p + ".conv.weight", C, 0);
                bf16* dUp = (bf16*)scratch(0, (size_t)rows2 * C * 2);
                dgrad(r.p + ".conv", {r.p + ".conv.weight"}, {dO}, {C}, {9}, H2, W2, C, dUp, nullptr);
                bf16* dX = (bf16*)scratch(1, (size_t)B * H * W * C * 2);
                op([=](cudaStream_t st) {
                    sumpool2(dUp, dX, Bn, H, W, C, st);
                    return (int)cudaGetLastError();
                });
                accumulate(r.xa, dX, C);
            } else if (r.kind == 3) {
                cur_label = "bwd " + r.p;
                const int C = r.xa.C, H = r.xa.H, W = r.xa.W, h = r.out.H, w = r.out.W;
                const bf16* dO = complete_grad(r.out);
                bias_grad(dO, (long long)B * h * w, C, r.p + ".conv.bias");
                bf16* dz = (bf16*)scratch(0, (size_t)B * H * W * C * 2);
                op([=](cudaStream_t st) {
                    zero_insert2x(dO, dz, Bn, h, w, C, st);
                    return (int)cudaGetLastError();
                });
                wgrad_sliced(dz, C, r.xa.p, C, H, W, C, C, 9, r.p + ".conv.weight", C, 0);
                bf16* dX = (bf16*)scratch(1, (size_t)B * H * W * C * 2);
                dgrad(r.p + ".conv", {r.p + ".conv.weight"}, {dz}, {C}, {9}, H, W, C, dX, nullptr);
                accumulate(r.xa, dX, C);
            } else if (r.kind == 0) {
                cur_label = "bwd conv_in";
                const bf16* dZ = complete_grad(r.out);
                bias_grad(dZ, (long long)B * R * R, ch, "conv_in.bias");
                float* wsf = (float*)scratch(4, (size_t)B * (R / 4) * ch * 27 * sizeof(float));
                float** g = gslot("conv_in.weight");
                op([=](cudaStream_t st) {
                    if (!*g) return 0;
                    conv_first_wgrad(dZ, pl->x, wsf, *g, Bn, R, R, ch, st);
                    return (int)cudaGetLastError();
                }, 2);
                {
                    // gradient w.r.t. the input state x (only when the caller asks for it: backward through a whole rollout,
                    // VARSampler.sample(enable_grad=True)): 3-row data-gradient GEMM of conv_in, fp32 NCHW output
                    dxmi_gemm_desc d = conv_desc(R, R);
                    set_src(d, 0, dZ, ch, ch);
                    add_seg(d, 0, 9);
                    d.b_ptr = packed_dgrad("conv_in", {"conv_in.weight"}, 3);
                    d.b_rows = 3;
                    d.b_ld = 9LL * ch;
                    d.out = (void*)16;  // patched per call
                    d.ldo = 3;
                    d.out_fp32 = 1;
                    d.out_nchw = 1;
                    d.block_n = 32;
                    if (!dry && !err) {
                        GemmOp g2;
                        int rr = prepare_gemm(d, &g2);
                        if (rr) {
                            err = rr;
                            engine_set_error("prepare_gemm(dx): %s", gemm_op_last_error());
                        } else if (g2.use_v2) {
                            fail("internal: dx GEMM must use the direct-store kernel");
                        } else {
                            op([g2, pl](cudaStream_t st) {
                                if (!pl->dx) return 0;
                                GemmOp g3 = g2;
                                g3.p.out = pl->dx;
                                return run_gemm(g3, st);
                            });
                        }
                    }
                }
            }
        }
        // ---- time-embedding path: d_tproj [B, TP] -> temb_proj, dense.1, dense.0
        cur_label = "bwd temb";
        float* d_st = (float*)alloc((size_t)B * temb_ch * sizeof(float));   // grad w.r.t. swish(temb)
        float* d_s1 = (float*)alloc((size_t)B * temb_ch * sizeof(float));   // grad w.r.t. swish(t1)
        {
            const int tc = temb_ch, tp = TP;
            float *temb_ = temb, *t1_ = t1, *te_ = te;
            // swish(temb) / swish(t1) once (the Linear-backward kernels would otherwise re-evaluate expf per output row)
            float* s_temb = (float*)alloc((size_t)B * temb_ch * sizeof(float));
            float* s_t1 = (float*)alloc((size_t)B * temb_ch * sizeof(float));
            op([=](cudaStream_t st) {
                silu_f32(temb_, s_temb, (long long)Bn * tc, st);
                silu_f32(t1_, s_t1, (long long)Bn * tc, st);
                return (int)cudaGetLastError();
            }, 2);
            {
                // every ResBlock's temb_proj backward in TWO launches (was 2 per block): the layers share their input swish(temb)
                LinearStack base{};
                std::vector<float**> gws, gbs;
                if ((int)rb.size() > LinearStack::MAX) fail("too many ResBlocks for the stacked temb_proj backward");
                base.n_layers = (int)rb.size();
                for (size_t i = 0; i < rb.size() && i < (size_t)LinearStack::MAX; ++i) {
                    base.off[i] = tp_offs[i];
                    base.W[i] = f32(rb[i] + ".temb_proj.weight");
                    gws.push_back(gslot(rb[i] + ".temb_proj.weight"));
                    gbs.push_back(gslot(rb[i] + ".temb_proj.bias"));
                }
                base.off[base.n_layers] = tp;
                op([=](cudaStream_t st) {
                    LinearStack ls = base;
                    for (int i = 0; i < ls.n_layers; ++i) {  // gradient destinations are bound per backward call
                        ls.dW[i] = *gws[i];
                        ls.db[i] = *gbs[i];
                    }
                    linear_stack_bwd_w(d_tproj, tp, s_temb, tc, ls, Bn, tc, st);
                    linear_stack_bwd_x(d_tproj, tp, ls, d_st, tc, Bn, tc, st);
                    return (int)cudaGetLastError();
                }, 2);
            }
            float **g1w = gslot("temb.dense.1.weight"), **g1b = gslot("temb.dense.1.bias");
            float **g0w = gslot("temb.dense.0.weight"), **g0b = gslot("temb.dense.0.bias");
            op([=](cudaStream_t st) {
                silu_bwd_mul(d_st, temb_, (long long)Bn * tc, st);            // -> grad w.r.t. temb
                linear_bwd_w(d_st, tc, s_t1, tc, 0, *g1w, *g1b, Bn, tc, tc, st);
                linear_bwd_x(d_st, tc, w1, d_s1, tc, Bn, tc, tc, 0, st);
                silu_bwd_mul(d_s1, t1_, (long long)Bn * tc, st);             // -> grad w.r.t. t1
                linear_bwd_w(d_s1, tc, te_, ch, 0, *g0w, *g0b, Bn, tc, ch, st);
                return (int)cudaGetLastError();
            }, 5);
        }
        emit_bwd = false;
    }
};

int build_unet_train_plan(Net& net, Plan& plan) {
    if (net.a.arch != DXMI_ARCH_DDPM_UNET) {
        engine_set_error("U-Net training plans exist for the DDPM U-Net only");
        return -28;
    }
    if (net.a.precision != 0) {
        engine_set_error("training plans run in bf16 mode only");
        return -27;
    }
    return build_two_pass<DdpmTrainBuilder>(net, plan);
}

}  // namespace dxmi
