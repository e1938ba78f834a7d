#pragma once
#include <functional>
#include "../../include/dxmi_b200.h"
#include "gemm_tc.cuh"

namespace dxmi {

struct GemmOp {
    ConvGemmParams p;
    int block_n, m_tiles, n_tiles, batch;
    int use_v2;  // 1: persistent kernel (gemm_tc2.cu), 2: cta_group::2 pair kernel (gemm_tc2p.cu)
    double flops;  // algorithmic 2*M*N*K of this launch
};

int prepare_gemm(const dxmi_gemm_desc& d, GemmOp* op);
int run_gemm(const GemmOp& op, cudaStream_t st);
void set_block_n_256(int v);
void set_small_map_bn(int v);
void set_shift3(int v);
void set_wave_bn(int v);
void set_s3_stages_max(int v);
void set_s3_m2(int v);
void set_lean_epi(int v);
int halo_option();
void set_dbg_mode(int v);
void set_gemm_version(int v);
void set_halo(int v);
void set_pair(int v);
void set_pair_min(int v);
int halo_tiles_per_image(int H, int W);
void set_dbg_times(void* p);
void set_time_gemms(int v);
void set_timing_dump(const char* path);
int gemm_timing_collect(double* ms_total, double* flops_total, long long* launches);
const char* gemm_op_last_error();
// "time_gemms" profiling (bench.py roofline legs): other tcgen05 launches join the GEMM list; HBM-bound families are kept
// per category (1 = GroupNorm finalize + apply, 2 = transition step) with their algorithmic bytes
int run_timed_tensor(double flops, int M, int N, int K, int batch, cudaStream_t st, const std::function<int()>& launch);
void run_timed_aux(int cat, double bytes, cudaStream_t st, const std::function<void()>& launch);
int aux_timing_collect(int cat, double* ms_total, double* bytes_total, long long* launches);

}  // namespace dxmi
