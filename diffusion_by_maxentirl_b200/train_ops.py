"""Trainer-side fused ops of the config-#4 iteration (second half) behind the reference trainer's own call
shapes:

* `running_cost(state, next_state, beta_next)`       - `DxMI_Trainer.get_running_cost` (trainer.py:163-169), differentiable
* `clip_grad_norm_(parameters, max_norm)`            - `torch.nn.utils.clip_grad_norm_` (trainer.py:324-325, :388)
* `FusedAdam(params_or_groups, lr=...)`              - `torch.optim.Adam` as built in train_cifar10.py:283-296 (two-LR groups)

Each is one to three kernel launches over ALL parameter tensors (a device table of pointers) instead of torch's per-op /
per-foreach-bucket launches; reductions run in a fixed order (deterministic).  CUDA only: no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib as L


class _RunningCost(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, next_state, beta_next):
        s = state.detach().contiguous().float()
        ns = next_state.detach().contiguous().float()
        b = beta_next.detach().to(device=s.device, dtype=torch.float32).contiguous()
        B, chw = s.shape[0], s[0].numel()
        rc = torch.empty(B, device=s.device)
        L.check(L.lib().dxmi_running_cost_fwd(L.ptr(s), L.ptr(ns), L.ptr(b), L.ptr(rc), B, chw, L.stream_ptr(s)), "running_cost_fwd")
        ctx.save_for_backward(s, ns, b)
        ctx.need = (state.requires_grad, next_state.requires_grad)
        ctx.shape = state.shape
        return rc

    @staticmethod
    def backward(ctx, g):
        s, ns, b = ctx.saved_tensors
        B, chw = s.shape[0], s[0].numel()
        g = g.detach().contiguous().float()
        ds = torch.empty_like(s) if ctx.need[0] else None
        dn = torch.empty_like(ns) if ctx.need[1] else None
        if ds is not None or dn is not None:
            L.check(L.lib().dxmi_running_cost_bwd(L.ptr(s), L.ptr(ns), L.ptr(b), L.ptr(g), L.ptr(dn), L.ptr(ds), B, chw, L.stream_ptr(s)),
                    "running_cost_bwd")
        return (ds.view(ctx.shape) if ds is not None else None), (dn.view(ctx.shape) if dn is not None else None), None


def running_cost(state, next_state, beta_next):
    """((next_state - state) ** 2 / (2 beta_next)).view(B, -1).mean(1) in one kernel (+ one for the backward).
    `beta_next` [B]: the reference's `extract(self.betas_for_q, n_timesteps - t - 1, state)` flattened."""
    if state.device.type != "cuda":
        raise RuntimeError("running_cost runs on CUDA only (no CPU fallback)")
    return _RunningCost.apply(state, next_state, beta_next.reshape(-1))


def trainer_get_running_cost(self, state, next_state, pred_mean, pred_std, t):
    """Drop-in for DxMI_Trainer.get_running_cost (same signature; `install()` binds it onto the reference's trainer class)."""
    t_reversed = (self.n_timesteps - t - 1)
    beta_next = self.betas_for_q.to(state.device)[t_reversed.to(state.device)]
    return running_cost(state, next_state, beta_next)


class _Table:
    """Device table of (param, grad, exp_avg, exp_avg_sq, numel, lr) rows + the chunk map; rebuilt when a pointer changes."""

    def __init__(self):
        self.sig = None

    def build(self, rows, device):
        """rows: list of (param, grad or None, m or None, v or None, lr)"""
        sig = tuple((p.data_ptr(), g.data_ptr() if g is not None else 0, m.data_ptr() if m is not None else 0,
                     v.data_ptr() if v is not None else 0, float(lr)) for p, g, m, v, lr in rows)
        if sig == self.sig:
            return
        chunk = L.lib().dxmi_opt_chunk_elems()
        arr = (L.OptTensor * len(rows))()
        ct, cf = [], []
        for i, (p, g, m, v, lr) in enumerate(rows):
            for t in (p, g, m, v):
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.device != device):
                    raise RuntimeError("fused optimizer ops need contiguous fp32 CUDA tensors on one device")
            arr[i].param, arr[i].grad = p.data_ptr(), (g.data_ptr() if g is not None else 0)
            arr[i].exp_avg, arr[i].exp_avg_sq = (m.data_ptr() if m is not None else 0), (v.data_ptr() if v is not None else 0)
            arr[i].numel, arr[i].lr = p.numel(), float(lr)
            for first in range(0, max(p.numel(), 1), chunk):
                ct.append(i)
                cf.append(first)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = raw.to(device)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=device)
        self.chunk_first = torch.tensor(cf, dtype=torch.int64, device=device)
        self.n_chunks = len(ct)
        self.partial = torch.empty(self.n_chunks, device=device)
        self.norm_coef = torch.zeros(2, device=device)
        self.sig = sig


_clip_tables = {}


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ (L2) over all gradients in 3 launches; returns the total norm (0-d device tensor)."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("fused clip_grad_norm_: L2 norm only (what the DxMI trainer uses)")
    if torch.is_tensor(parameters):
        parameters = [parameters]
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.tensor(0.0)
    dev = ps[0].device
    if dev.type != "cuda":
        raise RuntimeError("fused clip_grad_norm_ runs on CUDA only (no CPU fallback)")
    key = tuple(id(p) for p in ps)
    tab = _clip_tables.setdefault(key, _Table())
    if len(_clip_tables) > 16:
        _clip_tables.clear()
        _clip_tables[key] = tab
    tab.build([(p.data, p.grad, None, None, 0.0) for p in ps], dev)
    L.check(L.lib().dxmi_opt_grad_norm(L.ptr(tab.table), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_first), tab.n_chunks, float(max_norm),
                                      L.ptr(tab.partial), L.ptr(tab.norm_coef), 1, L.stream_ptr(dev)), "opt_grad_norm")
    return tab.norm_coef[0].clone()


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (weight_decay 0, amsgrad off) with ONE kernel launch per step over every parameter of every group.
    `step(max_norm=...)` additionally folds clip-by-global-norm into the same pass (2 + 1 launches): the gradient is scaled on
    the fly, which equals `clip_grad_norm_` followed by `Adam.step()`.  State (`exp_avg`, `exp_avg_sq`, `step`) uses torch's
    names, so `state_dict()` / `load_state_dict()` interchange with torch.optim.Adam checkpoints."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tab = _Table()
        self.last_grad_norm = None

    @torch.no_grad()
    def step(self, closure=None, max_norm=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        rows, dev = [], None
        g0 = self.param_groups[0]
        for g in self.param_groups:
            if g["betas"] != g0["betas"] or g["eps"] != g0["eps"]:
                raise NotImplementedError("FusedAdam: betas / eps must be the same in every group (only lr differs in DxMI)")
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                rows.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], g["lr"], st))
                dev = p.device
        if not rows:
            return loss
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA only (no CPU fallback)")
        steps = {int(r[5]["step"].item()) for r in rows}
        if len(steps) != 1:
            raise NotImplementedError("FusedAdam: all parameters must have taken the same number of steps")
        step = steps.pop()
        tab = self._tab
        tab.build([r[:5] for r in rows], dev)
        lib = L.lib()
        coef = None
        if max_norm is not None:
            L.check(lib.dxmi_opt_grad_norm(L.ptr(tab.table), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_first), tab.n_chunks, float(max_norm),
                                           L.ptr(tab.partial), L.ptr(tab.norm_coef), 0, L.stream_ptr(dev)), "opt_grad_norm")
            coef = tab.norm_coef
            self.last_grad_norm = tab.norm_coef[0]
        b1, b2 = g0["betas"]
        L.check(lib.dxmi_opt_adam_step(L.ptr(tab.table), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_first), tab.n_chunks, L.ptr(coef),
                                       float(b1), float(b2), float(g0["eps"]), step, 1 if coef is not None else 0, L.stream_ptr(dev)),
                "opt_adam_step")
        # the kernel wrote through raw pointers: bump the version counters so that the drop-in networks re-pack their bf16
        # operand copies on the next forward (native.NativeNet._ensure_handle) exactly as after a torch.optim step
        torch.autograd.graph.increment_version([r[0] for r in rows])
        return loss
