"""Drop-in for the reference's models/cm/unet.py::UNetModel (ADM / EDM U-Net, :523-790) on the B200 path."""
import torch

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.native import NativeNet


class UNetModel(NativeNet):
    """Same constructor, `forward(x, timesteps, y=None)` and `convert_to_fp16()` contract as the reference, same
    state_dict keys / shapes / dtypes (`input_blocks.{n}.0.in_layers.2.weight`, `...1.qkv.weight` [3C, C, 1], ...).
    The forward is one `dxmi_unet_forward` call: bf16 tcgen05 convolutions and attention with fp32 accumulation, fp32
    GroupNorm statistics / softmax / embedding MLP, fp32 NCHW in and out.  Built configuration family: dims=2,
    resblock_updown=True, legacy attention, dropout 0 (every DxMI EDM config)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False):
        if dims != 2 or not resblock_updown:
            raise NotImplementedError("B200 UNetModel supports dims=2, resblock_updown=True (all DxMI EDM configs)")
        if num_heads_upsample not in (-1, num_heads):
            raise NotImplementedError("num_heads_upsample != num_heads is not built")
        if any(float(m) != int(m) for m in channel_mult):
            raise NotImplementedError("fractional channel_mult (image_size 512) is not built")
        channel_mult = tuple(int(m) for m in channel_mult)
        d = L.ArchDesc()
        d.arch = L.ARCH_ADM_UNET
        d.resolution, d.in_channels, d.out_channels, d.ch = int(image_size), int(in_channels), int(out_channels), int(model_channels)
        d.n_levels = len(channel_mult)
        for i, m in enumerate(channel_mult):
            d.ch_mult[i] = m
        d.num_res_blocks = int(num_res_blocks)
        attention_resolutions = tuple(int(r) for r in attention_resolutions)
        d.n_attn = len(attention_resolutions)
        for i, r in enumerate(attention_resolutions):
            d.attn_resolutions[i] = r
        d.num_classes = int(num_classes) if num_classes is not None else 0
        d.num_head_channels = int(num_head_channels)
        d.num_heads = int(num_heads)
        d.use_scale_shift_norm = int(bool(use_scale_shift_norm))
        d.resblock_updown = int(bool(resblock_updown))
        super().__init__(d)
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.dropout_p = float(dropout)
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample

    def _torso_conv_params(self):
        for name, p in self.named_parameters():
            if name.split(".")[0] in ("input_blocks", "middle_block", "output_blocks"):
                stem = name.rsplit(".", 1)[0]
                w = self._param(stem + ".weight")
                if w.dim() >= 3:  # Conv1d / Conv2d only, like convert_module_to_f16 (models/cm/fp16_util.py:15-32)
                    yield p

    def convert_to_fp16(self):
        """Torso conv weights / biases -> fp16 (state_dict dtype contract of models/cm/unet.py:745-751); the packed
        bf16 operand copies are rebuilt from them on the next forward."""
        for p in self._torso_conv_params():
            p.data = p.data.half()

    def convert_to_fp32(self):
        for p in self._torso_conv_params():
            p.data = p.data.float()

    def forward(self, x, timesteps, y=None, x_scale=None):
        """x [N,C,H,W], timesteps [N] (the EDM `rescaled_t`), y [N] labels iff class-conditional.  `x_scale` ([N],
        optional) multiplies x per sample while it is loaded (the EDM c_in, karras_diffusion.py:349)."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        assert x.shape[2] == x.shape[3] == self.image_size
        self._check_eval()
        h = self._ensure_handle(x.device)
        B = x.shape[0]
        x = x.detach().contiguous().float()
        t = timesteps.detach().to(device=x.device, dtype=torch.float32).contiguous()
        assert t.shape == (B,)
        if y is not None:
            assert y.shape == (B,)
            y = y.detach().to(device=x.device, dtype=torch.long).contiguous()
        if x_scale is not None:
            x_scale = x_scale.detach().to(device=x.device, dtype=torch.float32).expand(B).contiguous()
        out = torch.empty(B, self.out_channels, self.image_size, self.image_size, device=x.device)
        L.check(L.lib().dxmi_unet_forward(h, L.ptr(x), L.ptr(x_scale), L.ptr(t), L.ptr(y), L.ptr(out), B, L.stream_ptr(x)),
                "dxmi_unet_forward")
        return out
