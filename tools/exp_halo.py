"""Experiment: shifted (non-1024-aligned) SWIZZLE_128B descriptors over a TMA halo tile."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

C.CDLL(L.LIB_PATH, mode=C.RTLD_GLOBAL)
lib = C.CDLL(os.path.join(ROOT, "tools", "libdxmi_exp.so"))  # make -C tools/csrc
lib.dxmi_exp_halo_conv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
torch.manual_seed(0)
H = 16
x = torch.randn(1, 64, H, 32, device="cuda")
xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
w = torch.randn(64, 64, 3, 3, device="cuda") / 24
wp = ops.pack_conv_weight(w)
ref = F.conv2d(xb.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1)[0]  # [64, H, 32]
for mode in (0, 1):
    for h0 in (0, 5):
        out = torch.zeros(128, 64, device="cuda")
        rc = lib.dxmi_exp_halo_conv(xb.data_ptr(), H, wp.data_ptr(), out.data_ptr(), h0, mode, None)
        torch.cuda.synchronize()
        errs = []
        for p in range(128):
            hh, ww = p // 34, p % 34
            if ww < 32 and h0 + hh < H:
                errs.append(float((out[p] - ref[:, h0 + hh, ww]).abs().max()))
        print(f"mode={mode} h0={h0} rc={rc} max|err| over {len(errs)} valid positions: {max(errs):.4g} (ref scale {float(ref.abs().mean()):.3f})")
