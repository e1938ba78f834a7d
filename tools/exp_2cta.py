"""Experiment: cta_group::2 GEMM mechanics."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402

C.CDLL(L.LIB_PATH, mode=C.RTLD_GLOBAL)
lib = C.CDLL(os.path.join(ROOT, "tools", "libdxmi_exp.so"))  # make -C tools/csrc
lib.dxmi_exp_2cta_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
torch.manual_seed(0)
for M, K in ((256, 64), (512, 256), (256 * 148, 256)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(128, K, device="cuda").to(torch.bfloat16)
    out = torch.zeros(M, 128, device="cuda")
    rc = lib.dxmi_exp_2cta_gemm(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, K, None)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = float((out - ref).abs().max())
    print(f"M={M} K={K} rc={rc} max|err|={err:.4g} (ref scale {float(ref.abs().mean()):.3f})", flush=True)
    if M == 512:
        for blk in range(4):
            e = float((out[blk * 128:(blk + 1) * 128] - ref[blk * 128:(blk + 1) * 128]).abs().max())
            print(f"   rows {blk*128}-{blk*128+127}: {e:.4g}; cols<64 {float((out[blk*128:(blk+1)*128,:64]-ref[blk*128:(blk+1)*128,:64]).abs().max()):.4g} "
                  f"cols>=64 {float((out[blk*128:(blk+1)*128,64:]-ref[blk*128:(blk+1)*128,64:]).abs().max()):.4g}")
