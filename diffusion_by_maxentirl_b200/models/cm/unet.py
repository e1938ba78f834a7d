"""Drop-in for the reference's models/cm/unet.py::UNetModel (ADM / EDM U-Net, :523-790) on the B200 path."""
import torch

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.native import NativeNet


class _AdmFunction(torch.autograd.Function):
    """F = net(x, timesteps, y) under autograd on the B200 path (trainer.py:693-746 update_sampler_mixed_precision: one
    `sample_step` per optimizer step): the forward keeps the activations in the handle's training plan
    (dxmi_adm_forward_train); the backward is one dxmi_unet_backward call that writes every parameter gradient in fp32 - they
    are returned in each parameter's own dtype (fp16 for the converted torso convolutions, like torch autograd would)."""

    @staticmethod
    def forward(ctx, module, x, t, y, *params):
        h = module._ensure_handle(x.device)
        B = x.shape[0]
        xc = x.detach().contiguous().float()
        tc = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        yc = y.detach().to(device=x.device, dtype=torch.long).contiguous() if y is not None else None
        out = torch.empty(B, module.out_channels, module.image_size, module.image_size, device=x.device)
        p_drop = float(module.dropout_p) if module.training else 0.0
        seed = int(torch.randint(0, 2**62, (1,)).item()) if p_drop > 0 else 0
        L.check(L.lib().dxmi_adm_forward_train(h, L.ptr(xc), L.ptr(tc), L.ptr(yc), L.ptr(out), p_drop, seed, B, L.stream_ptr(xc)),
                "dxmi_adm_forward_train")
        ctx.module, ctx.B, ctx.x = module, B, xc
        ctx.keep = (tc, yc)  # the plan reads the labels again in the backward (label_emb gradient): keep the buffers alive
        ctx.token = module._train_token = object()
        ctx.need_param = [p.requires_grad for p in params]
        ctx.param_dtypes = [p.dtype for p in params]
        if x.requires_grad:
            raise NotImplementedError("B200 ADM U-Net: the gradient w.r.t. the input state is not built (the EDM sampler update "
                                      "differentiates one step from a replay-buffer state, trainer.py:693-746)")
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.module
        h = m._ensure_handle(ctx.x.device)
        lib = L.lib()
        if m._train_token is not ctx.token:
            raise RuntimeError(
                "B200 U-Net: backward() of a forward whose saved activations were overwritten by a later grad-enabled forward at "
                "the same batch size (the plan keeps one set per batch size; run forward/backward pairs in order)")
        keys = m._keys
        sizes = [m._param(k).numel() for k in keys]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=ctx.x.device)
        grads, off = [], 0
        for k, n, need, dt in zip(keys, sizes, ctx.need_param, ctx.param_dtypes):
            g = flat[off:off + n]
            off += n
            L.check(lib.dxmi_bind_grad(h, k.encode(), L.ptr(g) if need else None), f"bind_grad {k}")
            grads.append((g.view(m._param(k).shape), dt) if need else None)
        d = dout.detach().contiguous().float()
        L.check(lib.dxmi_unet_backward(h, L.ptr(ctx.x), L.ptr(d), None, ctx.B, L.stream_ptr(ctx.x)), "dxmi_unet_backward")
        m._train_token = None
        for k in keys:
            lib.dxmi_bind_grad(h, k.encode(), None)
        return (None, None, None, None, *[None if e is None else e[0].to(e[1]) for e in grads])


class UNetModel(NativeNet):
    """Same constructor, `forward(x, timesteps, y=None)` and `convert_to_fp16()` contract as the reference, same
    state_dict keys / shapes / dtypes (`input_blocks.{n}.0.in_layers.2.weight`, `...1.qkv.weight` [3C, C, 1], ...).
    The forward is one `dxmi_unet_forward` call: bf16 tcgen05 convolutions and attention with fp32 accumulation, fp32
    GroupNorm statistics / softmax / embedding MLP, fp32 NCHW in and out.  In train() mode under autograd the forward records
    a graph (`_AdmFunction`: every parameter gradient from the B200 backward plan, engine_train_adm.cu).  Built configuration
    family: dims=2, resblock_updown=True, legacy attention (every DxMI EDM config)."""

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
        self._train_token = None

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
        if torch.is_grad_enabled() and self.training:
            # update_sampler_mixed_precision (trainer.py:693-746): backward through the U-Net
            if self.precision != "bf16":
                raise RuntimeError("B200 U-Net training path runs in bf16 mode only")
            if x_scale is not None:
                x = x * x_scale.to(x.device).reshape(-1, 1, 1, 1)
            assert timesteps.shape == (x.shape[0],)
            return _AdmFunction.apply(self, x, timesteps, y, *[self._param(k) for k in self._keys])
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
