"""Drop-in for the reference's models/modules.py: the energy / value network IGEBMEncoderV2 (:104-163) and
`process_single_t` (:183-186) on the B200 path."""
import torch

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.native import NativeNet


def process_single_t(x, t):
    """Broadcast a scalar step index to a [B] long tensor on x's device (reference modules.py:183-186)."""
    if isinstance(t, int) or t.dim() == 0 or len(t) == 1:
        t = torch.ones([x.shape[0]], dtype=torch.long, device=x.device) * t
    return t


class _ValueNetFunction(torch.autograd.Function):
    """`loss.backward()` through the value net on the B200 path (trainer.py:252-264, :320-326, :369-389): forward keeps the
    activations inside the handle's training plan (dxmi_value_forward_train), backward is one dxmi_value_backward call that
    writes every parameter gradient (tensor-core wgrad / dgrad kernels) and, if the input requires it, d loss / d x."""

    @staticmethod
    def forward(ctx, module, x, *params):
        h = module._ensure_handle(x.device)
        B = x.shape[0]
        xc = x.detach().contiguous().float()
        out = torch.empty(B, 1, device=x.device)
        L.check(L.lib().dxmi_value_forward_train(h, L.ptr(xc), L.ptr(out), B, L.stream_ptr(xc)), "dxmi_value_forward_train")
        ctx.module, ctx.B = module, B
        ctx.x = xc
        ctx.token = module._train_token = object()  # identifies the forward whose activations the plan holds
        ctx.need_dx = x.requires_grad
        ctx.need_param = [p.requires_grad for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.module
        if m._train_token is not ctx.token:
            raise RuntimeError(
                "B200 value net: backward() of a forward whose saved activations were overwritten by a later grad-enabled "
                "forward at the same batch size (the plan keeps one set per batch size; run forward/backward pairs in order)")
        h = m._ensure_handle(ctx.x.device)
        lib = L.lib()
        keys = m._keys
        sizes = [m._param(k).numel() for k in keys]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=ctx.x.device)
        grads, off = [], 0
        for k, n, need in zip(keys, sizes, ctx.need_param):
            g = flat[off:off + n]
            off += n
            L.check(lib.dxmi_bind_grad(h, k.encode(), L.ptr(g) if need else None), f"bind_grad {k}")
            grads.append(g.view(m._param(k).shape) if need else None)
        dx = torch.empty_like(ctx.x) if ctx.need_dx else None
        d = dout.detach().contiguous().float().view(-1)
        L.check(lib.dxmi_value_backward(h, L.ptr(ctx.x), L.ptr(d), L.ptr(dx) if dx is not None else None, ctx.B,
                                        L.stream_ptr(ctx.x)), "dxmi_value_backward")
        m._train_token = None
        for k in keys:
            lib.dxmi_bind_grad(h, k.encode(), None)
        grads = [g.to(m._param(k).dtype) if g is not None else None for g, k in zip(grads, keys)]
        return (None, dx, *grads)


class IGEBMEncoderV2(NativeNet):
    """conv3 -> lrelu -> 6 ResBlockV2 -> relu -> sum over HW -> Linear(2nh, 1) -> Linear(1, 1); `forward(x)` -> [B, 1].
    Only the configuration every DxMI YAML uses is built (no spectral norm, no class embedding, keepdim=False,
    linear output)."""

    def __init__(self, in_chan=3, out_chan=1, n_class=None, use_spectral_norm=False, keepdim=True,
                 out_activation="linear", avg_pool_dim=1, learn_out_scale=False, nh=128):
        if use_spectral_norm or n_class is not None or keepdim or out_activation != "linear" or out_chan != 1:
            raise NotImplementedError(
                "B200 IGEBMEncoderV2 supports use_spectral_norm=False, n_class=None, keepdim=False, "
                "out_activation='linear', out_chan=1 (the value-net settings of every DxMI config)")
        d = L.ArchDesc()
        d.arch = L.ARCH_IGEBM_V2
        d.in_channels, d.out_channels, d.ch = int(in_chan), int(out_chan), int(nh)
        d.resolution = 0  # set per call from the input
        d.learn_out_scale = int(bool(learn_out_scale))
        super().__init__(d)
        self.keepdim = keepdim
        self.learn_out_scale = learn_out_scale
        self.pre_activation = None
        self._train_token = None
        self._by_res = {}

    def load_pretrained(self, ckpt):
        """Reference modules.py:165-180: load the `conv1.*` and `blocks.*` entries of `ckpt['state_dict']` (keys carry a
        4-character `net.` prefix); the head (`linear`, `out_scale`) keeps its initialisation.  Like the reference's
        sub-module `load_state_dict` calls this is strict over conv1 / blocks."""
        own = {k: self._param(k) for k in self._keys if k.startswith("conv1.") or k.startswith("blocks.")}
        seen = set()
        with torch.no_grad():
            for k, v in ckpt["state_dict"].items():
                k_ = k[4:]
                if k_.startswith("conv1") or k_.startswith("blocks"):
                    if k_ not in own:
                        raise RuntimeError(f"load_pretrained: unexpected key {k!r}")
                    if tuple(own[k_].shape) != tuple(v.shape):
                        raise RuntimeError(f"load_pretrained: size mismatch for {k!r}: {tuple(v.shape)} vs {tuple(own[k_].shape)}")
                    own[k_].copy_(v)
                    seen.add(k_)
        missing = sorted(set(own) - seen)
        if missing:
            raise RuntimeError(f"load_pretrained: missing keys {missing[:4]}{'...' if len(missing) > 4 else ''}")

    def forward(self, input, y=None, out=None):
        """`out` (optional, not in the reference): a contiguous fp32 CUDA tensor of B elements the energies are written into
        (e.g. the energy slot of a packed gather buffer); inference only."""
        if y is not None:
            raise NotImplementedError("class-conditional value net is not used by the built configs")
        B, _, H, W = input.shape
        assert H == W and H % 8 == 0
        if self._desc.resolution != H:
            # the plan is resolution specific: (re)create the handle for this input size
            self.release()
            self._desc.resolution = H
        params = [self._param(k) for k in self._keys]
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in params)):
            if self.precision != "bf16":
                raise RuntimeError("B200 value net: the fp32 mode is inference-only (validation against the reference's fp32 "
                                   "path); call it under torch.no_grad(), or use set_precision('bf16') to train")
            out = _ValueNetFunction.apply(self, input, *params)
            self.pre_activation = out
            return out
        h = self._ensure_handle(input.device)
        x = input.detach().contiguous().float()
        if out is None:
            out = torch.empty(B, 1, device=x.device)
        else:
            assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == B and out.device == x.device
            out = out.view(B, 1)
        L.check(L.lib().dxmi_value_forward(h, L.ptr(x), L.ptr(out), B, L.stream_ptr(x)), "dxmi_value_forward")
        self.pre_activation = out
        return out
