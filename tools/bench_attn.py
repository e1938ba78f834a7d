"""Microbenchmark of the fused d=64 attention kernel (attn_tc.cu) at the ImageNet-64 geometries (B = 64 per GPU).
Usage (under gpurun): python tools/bench_attn.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import ops  # noqa: E402

torch.manual_seed(0)
for B, heads, seq in [(64, 6, 1024), (64, 9, 256), (64, 12, 64)]:
    Cc = heads * 64
    qk = (torch.randn(B, seq, 2 * Cc, device="cuda") * 1.5).to(torch.bfloat16)
    v = torch.randn(B, seq, Cc, device="cuda").to(torch.bfloat16)
    vt = v.transpose(1, 2).contiguous()
    for _ in range(3):
        out = ops.attention(qk, vt, heads, 0.125)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        out = ops.attention(qk, vt, heads, 0.125)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    fl = 4.0 * B * heads * seq * seq * 64
    q = qk[:2, :, :Cc].float().view(2, seq, heads, 64).transpose(1, 2)
    k = qk[:2, :, Cc:].float().view(2, seq, heads, 64).transpose(1, 2)
    vv = v[:2].float().view(2, seq, heads, 64).transpose(1, 2)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ vv).transpose(1, 2).reshape(2, seq, Cc)
    err = ((out[:2].float() - ref).norm() / ref.norm()).item()
    print(f"B={B} heads={heads} seq={seq}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  rel-L2 {err:.2e}")
