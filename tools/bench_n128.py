"""Where does a 32x32 128->128 3x3 conv (B=256: M=262144, N=128, K=1152) spend its time?  Pair kernel, shift3 on/off,
epilogue variants, dbg_mode 1 (no MMA) / 2 (no epilogue stores)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

dev = "cuda"
lib = L.lib()
N, H, C = 256, 32, 128
x = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, device=dev) / (3 * C**0.5)
b = torch.randn(C, device=dev)
wp = ops.pack_conv_weight(w)
M = N * H * H
out = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
res = torch.randn(M, C, device=dev).to(torch.bfloat16)
rv = torch.randn(N, C, device=dev)
st = torch.empty(M // 128, C, 2, device=dev)
flops = 2.0 * M * C * 9 * C


def run(tag, **kw):
    for _ in range(3):
        ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, bias=b, block_n=128, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, bias=b, block_n=128, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    print(f"  {tag:34s} {us:7.1f} us  {flops / us / 1e6:6.0f} TFLOP/s", flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "m2":
    for m2 in (0, 1):
        lib.dxmi_set_option(b"s3_m2", m2)
        print(f"two tiles per CTA (s3_m2) = {m2}")
        run("bias only")
        run("bias + stats", gn_stats=st, gn_seg=128)
        run("rowvec + stats (conv1)", rowvec=rv, gn_stats=st, gn_seg=128)
        run("residual + stats (conv2)", residual=res, gn_stats=st, gn_seg=128)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    for stg in (4, 3, 2):
        lib.dxmi_set_option(b"s3_stages_max", stg)
        print(f"shift3 stages <= {stg}")
        run("bias only")
        run("bias + stats", gn_stats=st, gn_seg=128)
        run("rowvec + stats (conv1)", rowvec=rv, gn_stats=st, gn_seg=128)
        run("residual + stats (conv2)", residual=res, gn_stats=st, gn_seg=128)
    sys.exit(0)
for s3 in (0, 1):
    lib.dxmi_set_option(b"shift3", s3)
    for dbg in (0, 1, 2):
        lib.dxmi_set_option(b"dbg_mode", dbg)
        print(f"shift3={s3} dbg_mode={dbg} ({['normal', 'no MMA', 'no epilogue stores'][dbg]})")
        run("bias only")
        run("bias + stats", gn_stats=st, gn_seg=128)
        run("rowvec + stats (conv1)", rowvec=rv, gn_stats=st, gn_seg=128)
        run("residual + stats (conv2)", residual=res, gn_stats=st, gn_seg=128)
lib.dxmi_set_option(b"dbg_mode", 0)
