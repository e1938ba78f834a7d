// Plan-builder base shared by the three network builders (engine.cu: DDPM U-Net + IGEBM value net, engine_adm.cu:
// ADM / EDM U-Net): bump-allocated activation arena (dry pass sizes it, real pass fills it), packed-weight cache and
// launch-closure emission.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "engine.cuh"

namespace dxmi {

int attnblk_option();   // engine.cu: 1 = DDPM AttnBlock at 16x16 as one fused kernel
int stats16_option();   // engine.cu
int conv_out_padded_option();
int up2_option();        // engine.cu: 1 = Upsample + conv3x3 as four phase convolutions of the low-resolution tensor
int gn_fused_option();  // engine.cu: 1 = every GroupNorm with producer statistics is ONE kernel (prologue + streaming apply)

struct Builder {
    Net& net;
    Plan& plan;
    bool dry;
    int B;
    size_t off = 0;
    static constexpr int NSLOT = 12;
    size_t scratch_max[NSLOT] = {0};
    size_t scratch_base[NSLOT] = {0};
    int err = 0;

    Builder(Net& n, Plan& p, bool d) : net(n), plan(p), dry(d), B(p.B) {}

    static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }
    void* alloc(size_t bytes) {
        void* r = dry ? nullptr : plan.arena + off;
        off += align_up(bytes);
        return r;
    }
    void* scratch(int slot, size_t bytes) {
        bytes = align_up(bytes);
        if (bytes > scratch_max[slot]) scratch_max[slot] = bytes;  // tracked in both passes (they must agree)
        if (dry) return nullptr;
        return plan.arena + scratch_base[slot];
    }
    bf16* act_alloc(int C, int H, int W) { return (bf16*)alloc((size_t)B * H * W * C * sizeof(bf16)); }
    // GroupNorm partial-statistics buffer for a GEMM output of B*H*W rows x C columns (null when the fused path
    // cannot be used for this geometry: a 32-row segment must not straddle two images)
    // rows per GroupNorm partial of a GEMM output.  16 (option "stats16", off by default - measured slower): a segment is exactly the 16 rows one
    // epilogue warp finishes, so the partial goes straight to global memory - no shared-memory combine and no second named
    // barrier per 32-column chunk in the epilogue (gemm_epi.cuh); the finalize then sums HW/16 partials per image.
    static int stats_seg(int HW) {
        if (stats16_option() && HW % 16 == 0) return 16;
        return HW % 128 == 0 ? 128 : (HW % 64 == 0 ? 64 : (HW % 32 == 0 ? 32 : (HW % 16 == 0 ? 16 : 0)));
    }
    static size_t stats_bytes(int rows, int HW, int C) { return (size_t)rows / stats_seg(HW) * C * 2 * sizeof(float); }
    // Layout of the GroupNorm partials a GEMM writes for its [B*H*W, C] output: per halo tile for 3x3 stride-1 convs on
    // 32- / 64-wide maps (gemm_op.cu: halo_tiles_per_image), else per 32/64/128-row segment.
    struct StatSpec {
        int P = 0;
        bool halo = false;
        size_t bytes = 0;
    };
    StatSpec stat_spec(int C, int H, int W, bool conv3x3_s1) const {
        StatSpec s;
        const int tpi = conv3x3_s1 ? halo_tiles_per_image(H, W) : 0;
        if (tpi) {
            s.P = tpi;
            s.halo = true;
        } else if (stats_seg(H * W)) {
            s.P = H * W / stats_seg(H * W);
        }
        s.bytes = (size_t)B * s.P * C * 2 * sizeof(float);
        return s;
    }
    // `conv3x3_s1`: the tensor will be produced by a 3x3 stride-1 convolution GEMM (its partials may be per halo tile)
    Act new_act(int C, int H, int W, bool want_stats = true, bool conv3x3_s1 = false) {
        Act a;
        a.p = act_alloc(C, H, W);
        a.C = C;
        a.H = H;
        a.W = W;
        const StatSpec s = stat_spec(C, H, W, conv3x3_s1);
        a.has_stats = want_stats && s.P > 0;
        if (a.has_stats) {
            a.stats = (float*)alloc(s.bytes);
            a.stats_P = s.P;
            a.stats_halo = s.halo;
        }
        return a;
    }
    // ---- nearest-2x upsample folded into the 3x3 convolution after it (gemm_tc.cuh ConvGemmParams::up2)
    bool up2_ok(int C, int H, int W) const {
        return up2_option() && !stats16_option() && C % 64 == 0 && (W & (W - 1)) == 0 && stats_seg(H * W) >= 32;
    }
    // partials per OUTPUT image the up2 GEMM writes: [4 phases][segments of the low-resolution image]
    static int up2_stats_P(int H, int W) { return 4 * (H * W / stats_seg(H * W)); }
    Act new_act_up2(int C, int H, int W) {
        Act a;
        a.p = act_alloc(C, 2 * H, 2 * W);
        a.C = C;
        a.H = 2 * H;
        a.W = 2 * W;
        a.has_stats = true;
        a.stats_P = up2_stats_P(H, W);
        a.stats_halo = false;
        a.stats = (float*)alloc((size_t)B * a.stats_P * C * 2 * sizeof(float));
        return a;
    }
    // out [B, 2H, 2W, Cout] = conv3x3(nearest_upsample2x(x [B, H, W, C])) + bias (+ per-image row vector); stats (optional): B * up2_stats_P partials
    void conv_up2(const bf16* x, int C, int H, int W, const std::string& wkey, const std::string& bkey, int Cout, bf16* out, float* stats,
                  const float* rowvec, int ldrv) {
        bf16* wp = nullptr;
        if (!dry) {
            const Bound* b = get(wkey);
            if (!b) return;
            if (b->shape.size() != 4 || b->shape[0] != Cout || b->shape[1] != C || b->shape[2] != 3 || b->shape[3] != 3) {
                fail("conv_up2: expected a [Cout, C, 3, 3] weight");
                return;
            }
            bool fresh = false;
            wp = (bf16*)derived_buf("wup2:" + wkey, (size_t)16 * Cout * C * sizeof(bf16), &fresh);
            if (!wp) return;
            if (fresh) {
                Net* np = &net;
                net.pack_jobs.push_back([np, wkey, Cout, C, wp](cudaStream_t st) {
                    const Bound& bb = np->bound[wkey];
                    pack_conv_weight_up2(bb.ptr, bb.dtype == DXMI_F16, Cout, C, wp, st);
                    count_launches(1);
                });
            }
        }
        dxmi_gemm_desc d = conv_desc(H, W);
        d.up2 = 1;
        set_src(d, 0, x, C, C);
        add_seg(d, 0, 4);
        d.batch = 4;
        d.b_batched = 1;
        d.b_ptr = wp;
        d.b_rows = Cout;
        d.b_ld = 4LL * C;
        d.b_batch_stride = (long long)Cout * 4 * C;
        d.bias = f32(bkey);
        d.rowvec = rowvec;
        d.ldrv = ldrv;
        d.out = out;
        d.ldo = Cout;
        d.gn_stats = stats;
        gemm(d);
    }

    static void want_stats(dxmi_gemm_desc& d, const Act& out) {
        d.gn_stats = out.stats;
        d.gn_halo_P = out.stats_halo ? out.stats_P : 0;
    }

    void fail(const char* what) {
        if (!err) {
            err = -20;
            engine_set_error("%s", what);
        }
    }

    // ---------------------------------------------------------------- weights
    const Bound* get(const std::string& key) {
        auto it = net.bound.find(key);
        if (it == net.bound.end() || !it->second.ptr) {
            if (!dry) {
                std::string m = "weight not bound: " + key;
                fail(m.c_str());
            }
            return nullptr;
        }
        return &it->second;
    }
    // fp32 view of a bound tensor: the borrowed pointer itself when fp32, else a converted (owned) copy.
    const float* f32(const std::string& key) {
        if (dry) return nullptr;
        const Bound* b = get(key);
        if (!b) return nullptr;
        if (b->dtype == DXMI_F32) return (const float*)b->ptr;
        auto it = net.derived.find("f32:" + key);
        if (it != net.derived.end()) return (const float*)it->second;
        long long n = 1;
        for (auto s : b->shape) n *= s;
        float* d = nullptr;
        cudaMalloc(&d, n * sizeof(float));
        net.owned.push_back(d);
        net.derived["f32:" + key] = d;
        Net* np = &net;
        net.pack_jobs.push_back([np, key, d, n](cudaStream_t st) {
            const Bound& bb = np->bound[key];
            cast_to_f32(bb.ptr, bb.dtype == DXMI_F16, d, n, st);
            count_launches(1);
        });
        return d;
    }
    void* derived_buf(const std::string& name, size_t bytes, bool* fresh) {
        auto it = net.derived.find(name);
        if (it != net.derived.end()) {
            *fresh = false;
            return it->second;
        }
        void* d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) {
            fail("cudaMalloc failed for packed weights");
            return nullptr;
        }
        net.owned.push_back(d);
        net.derived[name] = d;
        *fresh = true;
        return d;
    }
    struct PackPart {
        std::string key;  // conv weight key
        int c_off, c_cnt; // input-channel slice
        int row_off = 0, row_cnt = 0;  // output-row slice (row_cnt == 0: all rows)
    };
    // bf16 [rows, K] with K = sum over parts of taps*c_cnt; rows stacked from `row_parts` groups when stacking q|k|v.
    bf16* packed_rows(const std::string& name, const std::vector<std::vector<PackPart>>& row_groups, long long* K_out,
                      int* rows_out) {
        if (dry) return nullptr;
        // geometry
        long long K = 0;
        int rows = 0;
        for (size_t g = 0; g < row_groups.size(); ++g) {
            long long kg = 0;
            int rg = 0;
            for (auto& part : row_groups[g]) {
                const Bound* b = get(part.key);
                if (!b) return nullptr;
                const int taps = (int)(b->shape.size() == 4 ? b->shape[2] * b->shape[3] : 1);
                kg += (long long)taps * part.c_cnt;
                rg = part.row_cnt ? part.row_cnt : (int)b->shape[0];
            }
            if (g == 0) K = kg;
            if (kg != K) {
                fail("packed_rows: inconsistent K across row groups");
                return nullptr;
            }
            rows += rg;
        }
        if (K_out) *K_out = K;
        if (rows_out) *rows_out = rows;
        bool fresh = false;
        bf16* d = (bf16*)derived_buf("w:" + name, (size_t)rows * K * sizeof(bf16), &fresh);
        if (!d || !fresh) return d;
        Net* np = &net;
        int row0 = 0;
        for (auto& grp : row_groups) {
            long long k_off = 0;
            int rg = 0;
            for (auto& part : grp) {
                const Bound* b = get(part.key);
                const int kh = b->shape.size() == 4 ? (int)b->shape[2] : 1;
                const int kw = b->shape.size() == 4 ? (int)b->shape[3] : 1;
                const int Cin = (int)b->shape[1];
                const int Cout = part.row_cnt ? part.row_cnt : (int)b->shape[0];
                const long long w_off = (long long)part.row_off * Cin * kh * kw;  // elements into the OIHW tensor
                bf16* dst = d + (long long)row0 * K;
                const std::string key = part.key;
                const int c_off = part.c_off, c_cnt = part.c_cnt;
                net.pack_jobs.push_back([np, key, Cout, Cin, kh, kw, c_off, c_cnt, dst, K, k_off, w_off](cudaStream_t st) {
                    const Bound& bb = np->bound[key];
                    const bool half = bb.dtype == DXMI_F16;
                    const char* src = (const char*)bb.ptr + w_off * (half ? 2 : 4);
                    pack_conv_weight(src, half, Cout, Cin, kh, kw, c_off, c_cnt, dst, K, k_off, st);
                    count_launches(1);
                });
                k_off += (long long)kh * kw * c_cnt;
                rg = Cout;
            }
            row0 += rg;
        }
        return d;
    }
    // fp32 concatenation of several bound vectors / matrices (row-stacked)
    float* concat_f32(const std::string& name, const std::vector<std::string>& keys) {
        if (dry) return nullptr;
        long long total = 0;
        std::vector<long long> sizes;
        for (auto& k : keys) {
            const Bound* b = get(k);
            if (!b) return nullptr;
            long long n = 1;
            for (auto s : b->shape) n *= s;
            sizes.push_back(n);
            total += n;
        }
        bool fresh = false;
        float* d = (float*)derived_buf("cat:" + name, total * sizeof(float), &fresh);
        if (!d || !fresh) return d;
        Net* np = &net;
        long long o = 0;
        for (size_t i = 0; i < keys.size(); ++i) {
            const std::string key = keys[i];
            float* dst = d + o;
            const long long n = sizes[i];
            net.pack_jobs.push_back([np, key, dst, n](cudaStream_t st) {
                const Bound& bb = np->bound[key];
                cast_to_f32(bb.ptr, bb.dtype == DXMI_F16, dst, n, st);
                count_launches(1);
            });
            o += n;
        }
        return d;
    }
    float* sum_f32(const std::string& name, const std::string& ka, const std::string& kb, long long n) {
        if (dry) return nullptr;
        bool fresh = false;
        float* d = (float*)derived_buf("sum:" + name, n * sizeof(float), &fresh);
        if (!d || !fresh) return d;
        f32(ka);  // registers the fp16 -> fp32 cast jobs (if any) *before* this one, so ordering is correct
        f32(kb);
        Net* np = &net;
        // the operand pointers are resolved when the job runs: a borrowed parameter may have moved since (re-bind)
        net.pack_jobs.push_back([np, ka, kb, d, n](cudaStream_t st) {
            auto resolve = [np](const std::string& key) -> const float* {
                const Bound& bb = np->bound[key];
                if (bb.dtype == DXMI_F32) return (const float*)bb.ptr;
                return (const float*)np->derived["f32:" + key];
            };
            vec_add_f32(resolve(ka), resolve(kb), d, n, st);
            count_launches(1);
        });
        return d;
    }

    // ---------------------------------------------------------------- op emission
    std::string cur_label;  // set by the network builders before emitting a block
    bool emit_bwd = false;  // training plans: ops go to the backward list
    void op(std::function<int(cudaStream_t)> f, int launches = 1) {
        if (dry) return;
        if (emit_bwd) {
            plan.bwd_ops.push_back(std::move(f));
            plan.bwd_names.push_back(cur_label);
            plan.bwd_launches += launches;
            return;
        }
        plan.ops.push_back(std::move(f));
        plan.op_names.push_back(cur_label);
        plan.launches_per_run += launches;
    }
    void gemm(const dxmi_gemm_desc& d_in) {
        if (dry || err) return;
        dxmi_gemm_desc d = d_in;
        if (d.gn_stats) d.gn_seg = stats_seg(d.rows_per_image);
        GemmOp g;
        int r = prepare_gemm(d, &g);
        if (r) {
            err = r;
            engine_set_error("prepare_gemm: %s", gemm_op_last_error());
            return;
        }
        plan.gemm_flops += g.flops;
        const std::string keep = cur_label;
        char buf[160];
        snprintf(buf, sizeof buf, "%s gemm M=%d N=%d bn=%d v2=%d batch=%d", keep.c_str(), g.p.M_total, g.p.N_total, g.block_n, g.use_v2,
                 g.batch);
        cur_label = buf;
        op([g](cudaStream_t st) { return run_gemm(g, st); });
        cur_label = keep;
    }
    // GEMM whose fp32 NCHW output pointer is the per-call network output (plan.out)
    void gemm_to_plan_out(const dxmi_gemm_desc& d) {
        if (dry || err) return;
        GemmOp g;
        int r = prepare_gemm(d, &g);
        if (r) {
            err = r;
            engine_set_error("prepare_gemm: %s", gemm_op_last_error());
            return;
        }
        if (g.use_v2) {
            fail("internal: network-output GEMM must use the direct-store kernel");
            return;
        }
        plan.gemm_flops += g.flops;
        Plan* pl = &plan;
        op([g, pl](cudaStream_t st) {
            GemmOp g2 = g;
            g2.p.out = pl->out;
            return run_gemm(g2, st);
        });
    }
    // last 3x3 convolution of a network: NHWC bf16 -> fp32 NCHW [B, Cout, H, W] at plan.out
    void conv_out_nchw(const bf16* src, int C, int H, int W, const std::string& wkey, const std::string& bkey, int Cout) {
        const bool direct = conv_out_padded_option() == 1 && conv3x3_last_supported(H, W, C, Cout);
        if (direct || (conv_out_padded_option() && Cout <= 4 && (H * W) % 128 == 0 && (W == 32 || W == 64))) {
            // Cout = 3 on a 32- / 64-wide map: the persistent kernel (two-phase coalesced epilogue, 8-stage ring) with the weights
            // zero-padded to 8 output rows and a 32-column tile, into a padded NHWC fp32 buffer, then one small NHWC -> NCHW pass.
            const long long K = 9LL * C;
            bf16* wp = nullptr;
            float* bp = nullptr;
            if (!dry) {
                bool fresh = false, fresh_b = false;
                wp = (bf16*)derived_buf("w:pad8:" + wkey, (size_t)8 * K * sizeof(bf16), &fresh);
                bp = (float*)derived_buf("b:pad8:" + bkey, 8 * sizeof(float), &fresh_b);
                if (wp && bp && (fresh || fresh_b)) {
                    f32(bkey);
                    Net* np = &net;
                    net.pack_jobs.push_back([np, wkey, bkey, wp, bp, C, Cout, K](cudaStream_t st) {
                        const Bound& bw = np->bound[wkey];
                        const Bound& bb = np->bound[bkey];
                        cudaMemsetAsync(wp, 0, (size_t)8 * K * sizeof(bf16), st);
                        cudaMemsetAsync(bp, 0, 8 * sizeof(float), st);
                        pack_conv_weight(bw.ptr, bw.dtype == DXMI_F16, Cout, C, 3, 3, 0, C, wp, K, 0, st);
                        const float* bsrc = bb.dtype == DXMI_F32 ? (const float*)bb.ptr : (const float*)np->derived["f32:" + bkey];
                        cudaMemcpyAsync(bp, bsrc, (size_t)Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
                        count_launches(1);
                    });
                }
            }
            if (direct) {
                // read-once stream kernel (conv_last.cu): halo tile + weights in shared memory, mma.sync N = 8
                Plan* pl = &plan;
                ConvLastOp lop;
                if (!dry && prepare_conv3x3_last(src, B, H, W, C, &lop)) {
                    fail("conv3x3_last: tensor map encoding failed");
                    return;
                }
                op([=](cudaStream_t st) {
                    conv3x3_last(lop, wp, bp, (float*)pl->out, Cout, st);
                    return (int)cudaGetLastError();
                });
                return;
            }
            float* tmp = (float*)scratch(11, (size_t)B * H * W * 8 * sizeof(float));
            dxmi_gemm_desc d = conv_desc(H, W);
            set_src(d, 0, src, C, C);
            add_seg(d, 0, 9);
            d.b_ptr = wp;
            d.b_rows = 8;
            d.b_ld = K;
            d.bias = bp;
            d.out = tmp;
            d.ldo = 8;
            d.out_fp32 = 1;
            d.block_n = 32;
            // (measured: forcing the halo-tile mode here is slower still - 112 vs 74 us on the CIFAR map at B = 256; direct-store kernel: 92 us)
            gemm(d);
            Plan* pl = &plan;
            const int Bn = B, HW = H * W;
            op([=](cudaStream_t st) {
                nhwc_to_nchw_f32(tmp, 8, (float*)pl->out, Bn, HW, Cout, st);
                return (int)cudaGetLastError();
            });
            return;
        }
        dxmi_gemm_desc d = conv_desc(H, W);
        set_src(d, 0, src, C, C);
        add_seg(d, 0, 9);
        d.b_ptr = packed_rows(wkey, {{{wkey, 0, C}}}, nullptr, nullptr);
        d.b_rows = Cout;
        d.b_ld = 9LL * C;
        d.bias = f32(bkey);
        d.out = (void*)16;  // placeholder, patched per call
        d.ldo = Cout;
        d.out_fp32 = 1;
        d.out_nchw = 1;
        d.block_n = 32;
        gemm_to_plan_out(d);
    }
    // Y[B, O] (fp32) = silu(X[B, K]) . W^T + b for a row-stack of Linear layers (every ResBlock's time-embedding
    // projection in one tensor-core GEMM; unet_small.py:123, cm/unet.py:203-209)
    void batched_emb_projection(const float* x, int K, const std::string& name, const std::vector<std::string>& wkeys,
                                const std::vector<std::string>& bkeys, float* y, int O, int rows = -1) {
        const int Bn = rows > 0 ? rows : B;
        bf16* xb = (bf16*)alloc((size_t)Bn * K * sizeof(bf16));
        op([=](cudaStream_t st) {
            silu_to_bf16(x, xb, (long long)Bn * K, st);
            return (int)cudaGetLastError();
        });
        std::vector<std::vector<PackPart>> groups;
        for (auto& k : wkeys) groups.push_back({{k, 0, K}});
        dxmi_gemm_desc d;
        memset(&d, 0, sizeof d);
        d.N = 1;
        d.H = 1;
        d.W = Bn;
        d.out_H = 1;
        d.out_W = Bn;
        d.stride = 1;
        d.batch = 1;
        set_src(d, 0, xb, K, K);
        add_seg(d, 0, 1);
        d.b_ptr = packed_rows(name + ".weight", groups, nullptr, nullptr);
        d.b_rows = O;
        d.b_ld = K;
        d.bias = concat_f32(name + ".bias", bkeys);
        d.out = y;
        d.ldo = O;
        d.out_fp32 = 1;
        d.alpha = 1.f;
        d.rows_per_image = 1;
        if (K % 64 || O % 8) fail("batched_emb_projection: K must be a multiple of 64 and O of 8");
        if (Bn == 1 && O % 64 == 0) d.block_n = 64;  // one live row: more, narrower column tiles (same per-element accumulation order)
        gemm(d);
    }
    dxmi_gemm_desc conv_desc(int H, int W) {
        dxmi_gemm_desc d;
        memset(&d, 0, sizeof d);
        d.N = B;
        d.H = H;
        d.W = W;
        d.out_H = H;
        d.out_W = W;
        d.stride = 1;
        d.batch = 1;
        d.alpha = 1.f;
        d.rows_per_image = H * W;
        return d;
    }
    static void set_src(dxmi_gemm_desc& d, int i, const bf16* p, int C, int ld) {
        d.a_ptr[i] = p;
        d.a_C[i] = C;
        d.a_ld[i] = ld;
    }
    static void add_seg(dxmi_gemm_desc& d, int src, int taps) {
        d.seg_src[d.nseg] = src;
        d.seg_taps[d.nseg] = taps;
        d.nseg++;
    }

    // GroupNorm(32) over concat(x1, x2) -> out (scratch slot), optional SiLU / FiLM
    void group_norm(Act x1, Act x2, const std::string& pfx, float eps, int silu, const float* film, int film_ld,
                    bf16* out) {
        cur_label = "GN " + pfx;
        const int HW = x1.H * x1.W;
        const int slabs = gn_num_slabs(B, HW);
        float* ws = (float*)scratch(5, (size_t)B * slabs * 64 * sizeof(float));
        const float* gamma = f32(pfx + ".weight");
        const float* beta = f32(pfx + ".bias");
        const bf16 *p1 = x1.p, *p2 = x2.p;
        const int C1 = x1.C, C2 = x2.C;
        const int Bn = B;
        if ((C1 + C2) % 8 || (C1 % 8) || (C1 + C2) > 2048) fail("group_norm: unsupported channel count");
        const float* st1 = x1.stats;
        const float* st2 = x2.stats;
        const double gn_bytes = 4.0 * (double)B * HW * (C1 + C2);  // algorithmic: one bf16 read + one bf16 write per element
        if (x1.has_stats && (C2 == 0 || x2.has_stats)) {
            // statistics come from the producer GEMMs' epilogues: one pass over the tensor instead of two
            const int P1 = x1.stats_P, P2 = x2.stats_P;
            float* ab = (float*)scratch(7, (size_t)B * (C1 + C2) * 2 * sizeof(float));
            // option gn_fused: 0 = maps up to 8x8, 1 = every map, N > 1 = maps of up to N pixels
            const int fused_max_hw = gn_fused_option() == 1 ? (1 << 30) : (gn_fused_option() > 1 ? gn_fused_option() : 64);
            if (HW <= fused_max_hw && !x1.stats_halo && !x2.stats_halo) {
                // small maps (8x8, 4x4) are launch-latency bound: one kernel that derives the statistics in its prologue
                // (at most 2 partials per channel, one CTA per image) instead of finalize + apply
                op([=](cudaStream_t st) {
                    run_timed_aux(1, gn_bytes, st, [&] {
                        gn_apply_fused(p1, C1, C1, p2, C2, C2, Bn, HW, 32, eps, gamma, beta, film, film_ld, silu, st1, P1, st2, P2, out, st);
                    });
                    return (int)cudaGetLastError();
                });
                return;
            }
            op([=](cudaStream_t st) {
                run_timed_aux(1, gn_bytes, st, [&] {
                    gn_finalize_apply(p1, C1, C1, p2, C2, C2, Bn, HW, 32, eps, gamma, beta, film, film_ld, silu, st1, P1, st2, P2, ab, out, st);
                });
                return (int)cudaGetLastError();
            }, 2);
            return;
        }
        op([=](cudaStream_t st) {
            run_timed_aux(1, 1.5 * gn_bytes, st, [&] {  // two-pass form: the statistics pass reads the tensor once more
                gn_stats(p1, C1, C1, p2, C2, C2, Bn, HW, 32, ws, slabs, st);
                gn_apply(p1, C1, C1, p2, C2, C2, Bn, HW, 32, eps, gamma, beta, film, film_ld, silu, ws, slabs, out, st);
            });
            return (int)cudaGetLastError();
        },
           2);
    }
};


template <typename BuilderT>
int build_two_pass(Net& net, Plan& plan) {
    BuilderT dryb(net, plan, true);
    dryb.build();
    size_t total = dryb.off;
    size_t base[Builder::NSLOT];
    for (int s = 0; s < Builder::NSLOT; ++s) {
        base[s] = total;
        total += dryb.scratch_max[s];
    }
    plan.arena_bytes = total + 256;
    cudaError_t e = cudaMalloc((void**)&plan.arena, plan.arena_bytes);
    if (e != cudaSuccess) {
        engine_set_error("cudaMalloc(%zu bytes) for the B=%d activation arena failed: %s", plan.arena_bytes, plan.B,
                         cudaGetErrorString(e));
        return (int)e;
    }
    BuilderT b(net, plan, false);
    for (int s = 0; s < Builder::NSLOT; ++s) b.scratch_base[s] = base[s];
    b.build();
    if (b.err) return b.err;
    if (b.off != dryb.off) {
        engine_set_error("internal: dry/real arena mismatch (%zu vs %zu)", dryb.off, b.off);
        return -21;
    }
    for (int s = 0; s < Builder::NSLOT; ++s)
        if (b.scratch_max[s] > dryb.scratch_max[s]) {
            engine_set_error("internal: scratch slot %d grew between the sizing and the real pass (%zu vs %zu bytes)", s,
                             dryb.scratch_max[s], b.scratch_max[s]);
            return -23;
        }
    return 0;
}


}  // namespace dxmi
