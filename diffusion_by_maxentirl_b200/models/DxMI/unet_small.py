"""Drop-in for the reference's models/DxMI/unet_small.py::Model (DDPM U-Net, :194-332) on the B200 path."""
import torch

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.native import NativeNet


class _UNetFunction(torch.autograd.Function):
    """eps = net(x, t) under autograd on the B200 path (trainer.py:348-389, update_sampler): forward keeps the activations in
    the handle's training plan (dxmi_unet_forward_train), backward is one dxmi_unet_backward call that writes every parameter
    gradient and, when x requires grad (backward through a rollout), the gradient w.r.t. the input state."""

    @staticmethod
    def forward(ctx, module, x, t, recompute, *params):
        """recompute=False: the training plan keeps this forward's activations for ONE pending backward (update_sampler).
        recompute=True (activation checkpointing, used by VARSampler.sample(enable_grad=True) where T forwards are pending at
        once): only (x, t, dropout seed) are kept and the backward re-runs the training forward right before it walks the tape."""
        h = module._ensure_handle(x.device)
        B = x.shape[0]
        xc = x.detach().contiguous().float()
        tc = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty(B, module.out_ch, module.resolution, module.resolution, device=x.device)
        # training-mode dropout: a fresh seed per forward from torch's CPU generator (reproducible under torch.manual_seed); the
        # masks are counter-based functions of (seed, block, element) and are regenerated in the backward
        p_drop = float(module.dropout_p) if module.training else 0.0
        seed = int(torch.randint(0, 2**62, (1,)).item()) if p_drop > 0 else 0
        module._last_dropout_seed = seed
        if recompute and p_drop == 0.0:
            L.check(L.lib().dxmi_unet_forward(h, L.ptr(xc), None, L.ptr(tc), None, L.ptr(out), B, L.stream_ptr(xc)), "dxmi_unet_forward")
        else:
            L.check(L.lib().dxmi_unet_forward_train(h, L.ptr(xc), L.ptr(tc), L.ptr(out), p_drop, seed, B, L.stream_ptr(xc)),
                    "dxmi_unet_forward_train")
        ctx.module, ctx.B, ctx.x, ctx.t = module, B, xc, tc
        ctx.recompute, ctx.p_drop, ctx.seed = bool(recompute), p_drop, seed
        ctx.token = None
        if not recompute:
            ctx.token = module._train_token = object()
        ctx.need_param = [p.requires_grad for p in params]
        ctx.need_dx = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.module
        h = m._ensure_handle(ctx.x.device)
        lib = L.lib()
        if ctx.recompute:
            scratch = torch.empty(ctx.B, m.out_ch, m.resolution, m.resolution, device=ctx.x.device)
            L.check(lib.dxmi_unet_forward_train(h, L.ptr(ctx.x), L.ptr(ctx.t), L.ptr(scratch), ctx.p_drop, ctx.seed, ctx.B,
                                                L.stream_ptr(ctx.x)), "dxmi_unet_forward_train (recompute)")
            m._train_token = None
        elif m._train_token is not ctx.token:
            raise RuntimeError(
                "B200 U-Net: backward() of a forward whose saved activations were overwritten by a later grad-enabled forward at "
                "the same batch size (the plan keeps one set per batch size; run forward/backward pairs in order)")
        keys = m._keys
        sizes = [m._param(k).numel() for k in keys]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=ctx.x.device)
        grads, off = [], 0
        for k, n, need in zip(keys, sizes, ctx.need_param):
            g = flat[off:off + n]
            off += n
            L.check(lib.dxmi_bind_grad(h, k.encode(), L.ptr(g) if need else None), f"bind_grad {k}")
            grads.append(g.view(m._param(k).shape) if need else None)
        d = dout.detach().contiguous().float()
        dx = torch.empty_like(ctx.x) if ctx.need_dx else None
        L.check(lib.dxmi_unet_backward(h, L.ptr(ctx.x), L.ptr(d), L.ptr(dx), ctx.B, L.stream_ptr(ctx.x)), "dxmi_unet_backward")
        m._train_token = None
        for k in keys:
            lib.dxmi_bind_grad(h, k.encode(), None)
        return (None, dx, None, None, *grads)


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
        self._train_token = None
        self._last_dropout_seed = 0

    def dropout_streams(self):
        """{ResnetBlock prefix: stream id} of the training-mode dropout masks = the block's position on the forward tape of the
        training plan (conv_in = 0, then ResnetBlocks, AttnBlocks, Downsample / Upsample in execution order).  With
        `_last_dropout_seed` and `dxmi_op_dropout_mask` this reproduces the exact masks of the last training forward."""
        ids, pos = {}, 1
        d = self._desc
        attn = {d.attn_resolutions[i] for i in range(d.n_attn)}
        res = d.resolution
        for lvl in range(d.n_levels):
            for b in range(d.num_res_blocks):
                ids[f"down.{lvl}.block.{b}"] = pos
                pos += 2 if res in attn else 1
            if lvl != d.n_levels - 1:
                pos += 1
                res //= 2
        ids["mid.block_1"] = pos
        pos += 2
        ids["mid.block_2"] = pos
        pos += 1
        for lvl in reversed(range(d.n_levels)):
            for b in range(d.num_res_blocks + 1):
                ids[f"up.{lvl}.block.{b}"] = pos
                pos += 2 if res in attn else 1
            if lvl != 0:
                pos += 1
                res *= 2
        return ids

    def forward(self, x, t):
        assert x.shape[2] == x.shape[3] == self.resolution
        assert t.dim() == 1 and t.shape[0] == x.shape[0]
        ckpt = getattr(self, "_checkpoint_activations", False)
        if torch.is_grad_enabled() and (self.training or ckpt):
            # update_sampler (trainer.py:348-389): backward through the U-Net; sample(enable_grad=True): through a whole rollout
            if self.precision != "bf16":
                raise RuntimeError("B200 U-Net training path runs in bf16 mode only")
            return _UNetFunction.apply(self, x, t, ckpt, *[self._param(k) for k in self._keys])
        self._check_eval()
        h = self._ensure_handle(x.device)
        x = x.detach().contiguous().float()
        t = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty(x.shape[0], self.out_ch, self.resolution, self.resolution, device=x.device)
        L.check(L.lib().dxmi_unet_forward(h, L.ptr(x), None, L.ptr(t), None, L.ptr(out), x.shape[0], L.stream_ptr(x)),
                "dxmi_unet_forward")
        return out
