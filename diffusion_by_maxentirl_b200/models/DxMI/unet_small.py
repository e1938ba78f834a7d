"""Drop-in for the reference's models/DxMI/unet_small.py::Model (DDPM U-Net, :194-332) on the B200 path."""
import torch

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.native import NativeNet


class Model(NativeNet):
    """Same constructor and `forward(x, t)` contract as the reference; parameters carry the reference's state_dict keys
    (`temb.dense.0.weight`, `down.0.block.0.conv1.weight`, ...).  The forward is one `dxmi_unet_forward` call:
    tcgen05 implicit-GEMM convolutions / attention GEMMs in bf16 with fp32 accumulation, fp32 GroupNorm statistics,
    fp32 timestep-embedding MLP, fp32 NCHW in and out."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution):
        if not resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False is not used by any DxMI config and is not built")
        ch_mult = tuple(int(m) for m in ch_mult)
        attn_resolutions = tuple(int(r) for r in attn_resolutions)
        d = L.ArchDesc()
        d.arch = L.ARCH_DDPM_UNET
        d.resolution, d.in_channels, d.out_channels, d.ch = int(resolution), int(in_channels), int(out_ch), int(ch)
        d.n_levels = len(ch_mult)
        for i, m in enumerate(ch_mult):
            d.ch_mult[i] = m
        d.num_res_blocks = int(num_res_blocks)
        d.n_attn = len(attn_resolutions)
        for i, r in enumerate(attn_resolutions):
            d.attn_resolutions[i] = r
        super().__init__(d)
        self.ch = ch
        self.temb_ch = ch * 4
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.out_ch = out_ch
        self.dropout_p = float(dropout)

    def forward(self, x, t):
        assert x.shape[2] == x.shape[3] == self.resolution
        assert t.dim() == 1 and t.shape[0] == x.shape[0]
        self._check_eval()
        h = self._ensure_handle(x.device)
        x = x.detach().contiguous().float()
        t = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty(x.shape[0], self.out_ch, self.resolution, self.resolution, device=x.device)
        L.check(L.lib().dxmi_unet_forward(h, L.ptr(x), None, L.ptr(t), None, L.ptr(out), x.shape[0], L.stream_ptr()),
                "dxmi_unet_forward")
        return out
