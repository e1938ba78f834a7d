"""Phase timeline of the fused DDPM AttnBlock kernel (attnblk_tc.cu): globaltimer stamps per CTA. Usage: attnblk_timeline.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import build_ddpm  # noqa: E402
from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402

os.environ["DXMI_DBG_ATTNBLK"] = "1"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = L.lib()
dbg = torch.zeros(2 * B * 16, dtype=torch.int64, device="cuda")
lib.dxmi_set_debug_buffer(L.ptr(dbg))
net, sampler, value, sd, vsd = build_ddpm(4)
x = torch.randn(B, 3, 32, 32, device="cuda")
t = torch.full((B,), 100.0, device="cuda")
with torch.no_grad():
    for _ in range(3):
        net(x, t)
torch.cuda.synchronize()
d = dbg.view(2 * B, 16).cpu().double()  # stamps of the LAST attention block launch of the forward
lead = d[0::2]
t0 = lead[:, 0].min()
names = {0: "gn affine done", 1: "x landed", 2: "hn published", 3: "K acc done", 4: "Q drained", 5: "S done", 6: "O drained", 7: "Y done",
         8: "mma: hn ready", 9: "mma: K drained (T_A free)", 10: "mma: Q issued", 11: "mma: S issue", 12: "mma: O issue", 13: "mma: Y issue",
         14: "exit"}
order = [0, 1, 2, 8, 3, 9, 10, 4, 11, 5, 12, 6, 13, 7, 14]
start = lead[:, 0]
first = start < t0 + 3000  # clusters of the first wave
print(f"B={B}: kernel span {(d[:, 14].max() - t0) / 1e3:.1f} us; first-wave clusters {int(first.sum())}")
for sel, tag in ((first, "first wave"), (~first, "later waves")):
    if sel.sum() == 0:
        continue
    print(f"-- {tag}: mean time since the CTA's first stamp (us)")
    prev = 0.0
    for k in order:
        v = ((lead[sel, k] - lead[sel, 0]).mean() / 1e3).item()
        print(f"  {names[k]:28s} {v:8.2f}  (+{v - prev:.2f})")
        prev = v
raw = dbg.view(2 * B, 16).cpu()
for cta in (0, 1, 200, 201):
    r = raw[cta]
    base = int(raw[cta & ~1][0])
    print("cta", cta, [int(v) - base if int(v) else None for v in r.tolist()])
