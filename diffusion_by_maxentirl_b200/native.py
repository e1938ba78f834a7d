"""Base class of the drop-in networks: an ordinary `nn.Module` whose parameters carry the reference's `state_dict`
names, shapes and dtypes, and whose `forward` is one call into libdxmi_b200.so.

The module tree is generated from the key list the C library publishes for an architecture
(`dxmi_num_weights` / `dxmi_weight_key` / `dxmi_weight_shape`), so the Python side and the CUDA plan can never
disagree about the layout.  Parameters are *borrowed* by pointer: `load_state_dict`, `.to()`, `.half()`, optimizer
steps and DDP all keep working on plain tensors; before a forward the module re-binds pointers that moved and
re-packs weights whose version counter changed.
"""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib as L


class _Node(nn.Module):
    """Anonymous container; exists only so that parameter paths spell the reference's state_dict keys."""


def _init_like_torch(key, p):
    """Default PyTorch initialisers in the reference's registration order (Conv2d / Linear: kaiming_uniform(a=sqrt 5)
    for the weight then U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for the bias; GroupNorm: ones / zeros; Embedding: N(0,1))."""
    with torch.no_grad():
        if key.endswith("label_emb.weight"):
            p.normal_()
        elif key.endswith(".weight") and p.dim() >= 2:
            nn.init.kaiming_uniform_(p, a=math.sqrt(5))
        elif key.endswith(".weight"):
            p.fill_(1.0)
        else:
            p.zero_()


_PRECISIONS = {"bf16": 0, "fp32": 1}
_default_precision = os.environ.get("DXMI_PRECISION", "bf16")


def set_default_precision(mode):
    """Precision of networks constructed from now on: "bf16" (tcgen05 tensor-core path, rel-L2 <= 2e-2 against the
    reference) or "fp32" (CUDA-core FFMA validation mode, rel-L2 <= 1e-5 per step; DDPM U-Net and value net only).
    Also settable with the environment variable DXMI_PRECISION.  Not part of the reference's surface."""
    global _default_precision
    if mode not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {mode!r}")
    _default_precision = mode


class NativeNet(nn.Module):
    """nn.Module facade over a dxmi_net_t handle."""

    def __init__(self, desc):
        super().__init__()
        self._desc = desc
        if _default_precision not in _PRECISIONS:
            raise ValueError(f"DXMI_PRECISION must be one of {sorted(_PRECISIONS)}, got {_default_precision!r}")
        self._desc.precision = _PRECISIONS[_default_precision]
        self._handle = None
        self._handle_device = None
        self._bound_sig = None
        self._packed_version = None
        self._dirty = False
        self._keys = []
        self._build_parameters()

    # ------------------------------------------------------------------ parameter tree
    def _spec(self):
        lib = L.lib()
        h = C.c_void_p()
        L.check(lib.dxmi_create(C.byref(self._desc), 0, C.byref(h)), "dxmi_create")
        try:
            n = lib.dxmi_num_weights(h)
            buf = C.create_string_buffer(256)
            shape = (C.c_int64 * 8)()
            nd = C.c_int()
            out = []
            for i in range(n):
                lib.dxmi_weight_key(h, i, buf, 256)
                lib.dxmi_weight_shape(h, i, shape, C.byref(nd))
                out.append((buf.value.decode(), tuple(shape[j] for j in range(nd.value))))
            return out
        finally:
            lib.dxmi_destroy(h)

    def _build_parameters(self):
        spec = self._spec()
        pending_bias_bound = {}
        for key, shape in spec:
            parts = key.split(".")
            node = self
            for part in parts[:-1]:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            p = nn.Parameter(torch.empty(shape, dtype=torch.float32))
            _init_like_torch(key, p)
            if key.endswith(".weight") and p.dim() >= 2:
                fan_in = math.prod(shape[1:])
                pending_bias_bound[key[: -len(".weight")]] = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
            elif key.endswith(".bias") and key[: -len(".bias")] in pending_bias_bound:
                b = pending_bias_bound[key[: -len(".bias")]]
                with torch.no_grad():
                    p.uniform_(-b, b)
            node.register_parameter(parts[-1], p)
            self._keys.append(key)

    def _param(self, key):
        node = self
        parts = key.split(".")
        for part in parts[:-1]:
            node = node._modules[part]
        return node._parameters[parts[-1]]

    # ------------------------------------------------------------------ handle management
    def _ensure_handle(self, device):
        if device.type != "cuda":
            raise RuntimeError(
                "diffusion_by_maxentirl_b200 runs on CUDA (sm_100a) only: there is no CPU / PyTorch fallback for "
                f"{type(self).__name__}.forward (got device {device})"
            )
        lib = L.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != idx:
            self.release()
            h = C.c_void_p()
            L.check(lib.dxmi_create(C.byref(self._desc), idx, C.byref(h)), "dxmi_create")
            self._handle, self._handle_device = h, idx
            self._bound_sig = None
        params = [self._param(k) for k in self._keys]
        sig = tuple((p.data_ptr(), p.dtype) for p in params)
        version = sum(p._version for p in params)
        if sig != self._bound_sig:
            for k, p in zip(self._keys, params):
                if p.device.type != "cuda" or not p.is_contiguous():
                    raise RuntimeError(f"parameter {k} must be a contiguous CUDA tensor (is on {p.device})")
                if p.dtype == torch.float32:
                    dt = L.F32
                elif p.dtype == torch.float16:
                    dt = L.F16
                else:
                    raise RuntimeError(f"parameter {k}: dtype {p.dtype} not supported (fp32 / fp16 state_dicts only)")
                shape = (C.c_int64 * 8)(*p.shape)
                L.check(lib.dxmi_bind_weight(self._handle, k.encode(), L.ptr(p), dt, shape, p.dim()), f"bind {k}")
            L.check(lib.dxmi_finalize(self._handle, L.stream_ptr(device)), "dxmi_finalize")
            self._bound_sig, self._packed_version = sig, version
            self._dirty = False
        elif version != self._packed_version or self._dirty:
            L.check(lib.dxmi_repack(self._handle, L.stream_ptr(device)), "dxmi_repack")
            self._packed_version = version
            self._dirty = False
        return self._handle

    def mark_dirty(self):
        """Tell the module its parameters were modified in a way the version counters cannot see - in-place edits through
        `.data` (`p.data.copy_()`, `p.data.mul_()`, hand-written EMA / weight surgery).  The packed bf16 operand copies are
        refreshed on the next forward.  Optimizer steps, `load_state_dict`, `.to()` / `.half()` are detected automatically."""
        self._dirty = True
        return self

    def repack(self):
        """`mark_dirty()` + refresh the packed weights now (on the current stream of the parameters' device)."""
        self._dirty = True
        p = self._param(self._keys[0])
        if p.device.type == "cuda":
            self._ensure_handle(p.device)
        return self

    def set_precision(self, mode):
        """Switch this network between the "bf16" and "fp32" paths (the handle and its plans are rebuilt lazily)."""
        if mode not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {mode!r}")
        if self._desc.precision != _PRECISIONS[mode]:
            self.release()
            self._desc.precision = _PRECISIONS[mode]
        return self

    @property
    def precision(self):
        return "fp32" if self._desc.precision == 1 else "bf16"

    def release(self):
        if self._handle is not None:
            L.lib().dxmi_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _check_eval(self):
        """Guards of the inference entry points (the DDPM `Model.forward` handles train() mode under autograd itself)."""
        if self.training and getattr(self, "dropout_p", 0.0) > 0:
            raise RuntimeError(
                "the B200 inference path has no dropout: a train()-mode forward without autograd (where the reference would apply "
                "dropout) is not built - call .eval() for sampling (generate_*.py, trainer.py:179 do), or run under autograd for "
                "the sampler update"
            )
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError(
                "B200 net: this entry point does not record a graph (the U-Net `forward`s handle train() mode under autograd "
                "themselves). Use .eval() / torch.no_grad() for sampling."
            )
