// Tensor-core weight-gradient op (wgrad_tc.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace dxmi {

struct WgradOp {
    alignas(64) unsigned char params[512];  // WgradParams (two TMA descriptors + geometry)
    int S, Cout, Cin, taps, grid, smem;
    size_t partial_floats;  // workspace the launch needs: S * Cout * taps * Cin floats
    double flops;
};

// dy: NHWC bf16 [N,H,W,Cout] (gradient of the conv output), x: NHWC bf16 [N,H,W,Cin] (the conv input).
// dy_ld / x_ld: pixel strides in elements (0 = dense: Cout / Cin) - channel slices of wider tensors
int prepare_wgrad(const void* dy, const void* x, int N, int H, int W, int Cout, int Cin, int taps, WgradOp* op, long long dy_ld = 0,
                  long long x_ld = 0);
// grad: fp32 OIHW [Cout][Cin_total][k][k]; this op fills input channels ci_off .. ci_off + Cin.  grad = scale * dW.
int run_wgrad(const WgradOp& op, float* partial_ws, float* grad, int Cin_total, int ci_off, float scale, cudaStream_t st);

}  // namespace dxmi
