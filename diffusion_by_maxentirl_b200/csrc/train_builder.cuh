// Shared machinery of the U-Net training plans (engine_train_unet.cu: DDPM, engine_train_adm.cu: ADM / EDM): activation identities
// and their bf16 gradient buffers, the forward GroupNorm that keeps what the backward needs, and the backward operators
// (data gradients through the forward implicit-GEMM kernels on transposed weights, tcgen05 weight gradients, bias gradients,
// GroupNorm backward).  See engine_train_unet.cu for the reverse-mode structure.
#pragma once
#include <map>

#include "builder.cuh"
#include "kernels_bwd.cuh"
#include "wgrad_tc.cuh"

namespace dxmi {

struct TrainBuilder : Builder {
    using Builder::Builder;

    struct GnSave {
        float* ab = nullptr;  // [B][C][2]
        float* mr = nullptr;  // [B][32][2]
    };
    std::map<int, std::pair<bf16*, bool>> gbuf;  // activation id -> (gradient buffer, initialised)
    int next_id = 0;
    float gn_eps = 1e-6f;  // unet_small.py Normalize: eps 1e-6; ADM GroupNorm32: 1e-5
    Act mk(int C, int H, int W, bool stats = true) {
        Act a = new_act(C, H, W, stats, false);
        a.id = next_id++;
        return a;
    }

    float** gslot(const std::string& key) { return &net.grad[key]; }

    // ---------------------------------------------------------------- gradient buffers
    bf16* grad_of(const Act& a) {
        if (a.id < 0) fail("internal: gradient requested for an activation without identity");
        auto it = gbuf.find(a.id);
        if (it != gbuf.end()) return it->second.first;
        bf16* g = (bf16*)alloc((size_t)B * a.H * a.W * a.C * sizeof(bf16));
        gbuf[a.id] = {g, false};
        return g;
    }
    // grad(a) (+)= src[rows, 0:a.C] with row stride ld
    void accumulate(const Act& a, const bf16* src, long long ld) {
        bf16* g = grad_of(a);
        auto& e = gbuf[a.id];
        const int init = e.second ? 0 : 1;
        e.second = true;
        const long long rows = (long long)B * a.H * a.W;
        const int C = a.C;
        op([=](cudaStream_t st) {
            accum_bf16(g, src, ld, rows, C, init, st);
            return (int)cudaGetLastError();
        });
    }
    const bf16* complete_grad(const Act& a) {
        auto it = gbuf.find(a.id);
        if (it == gbuf.end() || !it->second.second) {
            fail("internal: activation without gradient contributions in the backward walk");
            return grad_of(a);
        }
        return it->second.first;
    }

    // ---------------------------------------------------------------- forward helpers
    GnSave gn_fwd(Act x1, Act x2, const std::string& pfx, int silu, bf16* out, const float* film = nullptr, int film_ld = 0) {
        cur_label = "GN " + pfx;
        GnSave s;
        const int C1 = x1.C, C2 = x2.C, HW = x1.H * x1.W;
        s.ab = (float*)alloc((size_t)B * (C1 + C2) * 2 * sizeof(float));
        s.mr = (float*)alloc((size_t)B * 32 * 2 * sizeof(float));
        if (!x1.has_stats || (C2 && !x2.has_stats) || x1.stats_halo || x2.stats_halo) fail("training GroupNorm needs producer statistics");
        const float* gamma = f32(pfx + ".weight");
        const float* beta = f32(pfx + ".bias");
        const bf16 *p1 = x1.p, *p2 = x2.p;
        const float *st1 = x1.stats, *st2 = x2.stats;
        const int P1 = x1.stats_P, P2 = x2.stats_P, Bn = B;
        float *ab = s.ab, *mr = s.mr;
        const float eps = gn_eps;
        op([=](cudaStream_t st) {
            gn_finalize_apply(p1, C1, C1, p2, C2, C2, Bn, HW, 32, eps, gamma, beta, film, film_ld, silu, st1, P1, st2, P2, ab, out, st, mr);
            return (int)cudaGetLastError();
        }, 2);
        return s;
    }
    Act conv3(const std::string& key, const bf16* src, int Cin, int H, int W, int Cout, const float* rowvec, int ldrv,
              const bf16* residual, bool stats = true) {
        Act out = mk(Cout, H, W, stats);
        dxmi_gemm_desc d = conv_desc(H, W);
        set_src(d, 0, src, Cin, Cin);
        add_seg(d, 0, 9);
        d.b_ptr = packed_rows(key, {{{key + ".weight", 0, Cin}}}, nullptr, nullptr);
        d.b_rows = Cout;
        d.b_ld = 9LL * Cin;
        d.bias = f32(key + ".bias");
        d.rowvec = rowvec;
        d.ldrv = ldrv;
        d.residual = residual;
        d.ldr = Cout;
        d.out = out.p;
        d.ldo = Cout;
        if (stats) want_stats(d, out);
        gemm(d);
        return out;
    }
    // batched GEMM per image: out[b] (M x N) = alpha * A[b] (M x K, row stride a_ld) . B[b] (N x K, row stride b_ld)^T
    void bgemm(const bf16* A, int K, int a_ld, int M, const bf16* Bm, int N, long long b_ld, long long b_bs, void* out, int ldo,
               long long out_bs, bool fp32, float alpha, bool softmax) {
        dxmi_gemm_desc d;
        memset(&d, 0, sizeof d);
        d.N = B;
        d.H = 1;
        d.W = M;
        d.out_H = 1;
        d.out_W = M;
        d.stride = 1;
        set_src(d, 0, A, K, a_ld);
        add_seg(d, 0, 1);
        d.a_batched = 1;
        d.b_ptr = Bm;
        d.b_rows = N;
        d.b_ld = b_ld;
        d.b_batch_stride = b_bs;
        d.b_batched = 1;
        d.batch = B;
        d.alpha = alpha;
        d.softmax = softmax ? 1 : 0;
        d.out = out;
        d.ldo = ldo;
        d.out_batch_stride = out_bs;
        d.out_fp32 = fp32 ? 1 : 0;
        d.rows_per_image = 1;
        gemm(d);
    }

    // ---------------------------------------------------------------- backward helpers
    bf16* packed_dgrad(const std::string& name, const std::vector<std::string>& keys, int rows) {
        if (dry) return nullptr;
        long long K = 0;
        for (auto& k : keys) {
            const Bound* b = get(k);
            if (!b) return nullptr;
            K += b->shape[0] * (b->shape.size() == 4 ? b->shape[2] * b->shape[3] : 1);  // (Conv1d k=1 weights are [O, I, 1])
        }
        bool fresh = false;
        bf16* d = (bf16*)derived_buf("wT:" + name, (size_t)rows * K * sizeof(bf16), &fresh);
        if (!d || !fresh) return d;
        Net* np = &net;
        long long k_off = 0;
        for (auto& k : keys) {
            const Bound* b = get(k);
            const int Cout = (int)b->shape[0], Cin = (int)b->shape[1], taps = b->shape.size() == 4 ? (int)(b->shape[2] * b->shape[3]) : 1;
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
    // grad[wkey][:, ci_off : ci_off + Cin] = wgrad(dy [rows, Cout] (stride dy_ld), x [rows, Cin] (stride x_ld))
    void wgrad(const bf16* dy, long long dy_ld, const bf16* x, long long x_ld, int H, int W, int Cout, int Cin, int taps,
               const std::string& wkey, int Cin_total, int ci_off) {
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
        op([w, ws, g, Cin_total, ci_off](cudaStream_t st) {
            if (!*g) return 0;
            return run_wgrad(w, ws, *g, Cin_total, ci_off, 1.f, st);
        }, 2);
    }
    // weight gradient of a conv whose input has Cin channels (row stride x_ld), in slices of at most 256 channels
    void wgrad_sliced(const bf16* dy, long long dy_ld, const bf16* x, long long x_ld, int H, int W, int Cout, int Cin, int taps,
                      const std::string& wkey, int Cin_total, int ci_off) {
        for (int c0 = 0; c0 < Cin;) {
            int c = Cin - c0;
            if (c > 256) c = 256;
            wgrad(dy, dy_ld, x + c0, x_ld, H, W, Cout, c, taps, wkey, Cin_total, ci_off + c0);
            c0 += c;
        }
    }
    void bias_grad(const bf16* dy, long long rows, int C, const std::string& bkey, const std::string& bkey2 = "") {
        float* ws = (float*)scratch(5, (size_t)colsum_ws_floats(rows, C) * sizeof(float));
        float* tmp = (float*)alloc((size_t)C * sizeof(float));
        float** g = gslot(bkey);
        float** g2 = bkey2.empty() ? nullptr : gslot(bkey2);
        op([=](cudaStream_t st) {
            if (!*g && !(g2 && *g2)) return 0;
            colsum_bf16(dy, rows, C, ws, tmp, st);
            if (*g) cudaMemcpyAsync(*g, tmp, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, st);
            if (g2 && *g2) cudaMemcpyAsync(*g2, tmp, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, st);
            return (int)cudaGetLastError();
        }, 2);
    }
    // dX [rows, Cin_rows] = sum over segments conv^T(src_i, W_i)  (+ residual)
    bf16* dgrad(const std::string& name, const std::vector<std::string>& keys, const std::vector<const bf16*>& srcs,
                const std::vector<int>& src_C, const std::vector<int>& taps, int H, int W, int rows_out, bf16* out, const bf16* residual) {
        dxmi_gemm_desc d = conv_desc(H, W);
        long long K = 0;
        for (size_t i = 0; i < srcs.size(); ++i) {
            set_src(d, (int)i, srcs[i], src_C[i], src_C[i]);
            add_seg(d, (int)i, taps[i]);
            K += (long long)taps[i] * src_C[i];
        }
        d.b_ptr = packed_dgrad(name, keys, rows_out);
        d.b_rows = rows_out;
        d.b_ld = K;
        d.residual = residual;
        d.ldr = rows_out;
        d.out = out;
        d.ldo = rows_out;
        gemm(d);
        return out;
    }
    void gn_bwd(const std::string& pfx, Act x1, Act x2, const bf16* dy, const GnSave& s, int silu, bf16* dx) {
        const int C = x1.C + x2.C, HW = x1.H * x1.W, Bn = B;
        float* ws = (float*)scratch(6, (size_t)gn_bwd_ws_floats(B, HW, C) * sizeof(float));
        float** gg = gslot(pfx + ".weight");
        float** gb = gslot(pfx + ".bias");
        const bf16 *p1 = x1.p, *p2 = x2.p;
        const int C1 = x1.C, C2 = x2.C;
        const float *ab = s.ab, *mr = s.mr;
        op([=](cudaStream_t st) {
            group_norm_bwd(p1, C1, p2, C2, dy, ab, mr, Bn, HW, 32, silu, ws, dx, *gg, *gb, st);
            return (int)cudaGetLastError();
        }, 3);
    }
};

}  // namespace dxmi
