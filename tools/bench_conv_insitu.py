"""Why are the 32x32 N=128 convs slower inside the network than in isolation?  Same GEMM with (a) nothing, (b) GroupNorm
partials, (c) + per-image row vector, (d) the same but rotating over 4 input/output buffer sets (cold L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import ops  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402

L.lib().dxmi_set_option(b"pair_resident_b", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
dev = "cuda"
N, H, Cin, Cout = 256, 32, 128, 128
w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
b = torch.randn(Cout, device=dev)
wp = ops.pack_conv_weight(w)
xs = [torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16) for _ in range(4)]
outs = [torch.empty(N * H * H, Cout, dtype=torch.bfloat16, device=dev) for _ in range(4)]
stats = torch.zeros(N * H * H // 128, Cout, 2, device=dev)
rv = torch.randn(N, Cout, device=dev)
flops = 2.0 * N * H * H * Cout * Cin * 9


def run(label, rot, **kw):
    def once(i):
        j = i % 4 if rot else 0
        ops.conv_gemm([(xs[j], Cin, Cin)], [(0, 9)], wp, N, H, H, bias=b, out=outs[j], **kw)

    for i in range(4):
        once(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        once(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{label:50s} {us:7.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


run("plain, warm L2", False)
run("+ GN partials (seg 128), warm L2", False, gn_stats=stats, gn_seg=128)
run("+ GN partials + row vector, warm L2", False, gn_stats=stats, gn_seg=128, rowvec=rv)
run("plain, rotating 4 buffer sets (cold L2)", True)
run("+ GN partials + row vector, cold L2", True, gn_stats=stats, gn_seg=128, rowvec=rv)
res = torch.randn(N * H * H, Cout, device=dev).to(torch.bfloat16)
run("+ residual + GN partials, warm L2", False, gn_stats=stats, gn_seg=128, residual=res)
