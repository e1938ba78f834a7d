"""Per-shape throughput of the tcgen05 implicit-GEMM conv kernel (CUDA events, back-to-back launches).
Usage (under gpurun): python tools/bench_conv.py [--set cifar|in64|all] [--iters 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--set", default="cifar")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--block-n", type=int, nargs="*", default=[0])
ap.add_argument("--dbg", type=int, default=0)
ap.add_argument("--gemm-version", type=int, default=2)
ap.add_argument("--halo", type=int, default=0)
ap.add_argument("--pair", type=int, default=0)
ap.add_argument("--resb", type=int, default=0)
args = ap.parse_args()

# (name, N, H, Cin, Cout, taps)
SHAPES = {
    "halo": [("c32_128_128", 256, 32, 128, 128, 9), ("c32_256_256", 256, 32, 256, 256, 9), ("i64_192_192", 64, 64, 192, 192, 9),
             ("i32_384_384", 64, 32, 384, 384, 9)],
    "cifar": [("c32_128_128", 256, 32, 128, 128, 9), ("c32_256_128", 256, 32, 256, 128, 9),
              ("c16_256_256", 256, 16, 256, 256, 9), ("c16_512_256", 256, 16, 512, 256, 9),
              ("c8_256_256", 256, 8, 256, 256, 9), ("c4_256_256", 256, 4, 256, 256, 9),
              ("p16_256_256_1x1", 256, 16, 256, 256, 1), ("p16_256_512_1x1", 256, 16, 256, 512, 1)],
    "in64": [("i64_192_192", 64, 64, 192, 192, 9), ("i32_384_384", 64, 32, 384, 384, 9),
             ("i16_576_576", 64, 16, 576, 576, 9), ("i8_768_768", 64, 8, 768, 768, 9),
             ("i32_768_384", 64, 32, 768, 384, 9), ("i32_384_1152_1x1", 64, 32, 384, 1152, 1)],
}
from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
L.lib().dxmi_set_option(b"dbg_mode", args.dbg)
L.lib().dxmi_set_option(b"gemm_version", args.gemm_version)
L.lib().dxmi_set_option(b"halo", args.halo)
L.lib().dxmi_set_option(b"pair", args.pair)
L.lib().dxmi_set_option(b"pair_resident_b", args.resb)
sets = ["cifar", "in64"] if args.set == "all" else [args.set]
dev = "cuda"
for s in sets:
    for name, N, H, Cin, Cout, taps in SHAPES[s]:
        x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
        k = 3 if taps == 9 else 1
        w = torch.randn(Cout, Cin, k, k, device=dev) / (k * Cin**0.5)
        b = torch.randn(Cout, device=dev)
        wp = ops.pack_conv_weight(w)
        out = torch.empty(N * H * H, Cout, dtype=torch.bfloat16, device=dev)
        flops = 2.0 * N * H * H * Cout * Cin * taps
        for bn in args.block_n:
            try:
                for _ in range(3):
                    ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, bias=b, block_n=bn, out=out)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, bias=b, block_n=bn, out=out)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / args.iters
                print(f"{name:22s} block_n={bn:3d}  {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"{name:22s} block_n={bn:3d}  FAILED {str(ex)[:100]}", flush=True)
