"""Host side of the reference's 16-bit training helper (models/cm/fp16_util.py) for the B200 path: the contract
`train_image_large.py:155-169,259-263` and `DxMI_Trainer.update_sampler_mixed_precision` (trainer.py:693-746) rely on.

* `convert_module_to_f16 / _f32` (fp16_util.py:15-32): Conv{1,2,3}d parameters only.
* `MixedPrecisionTrainer` (fp16_util.py:161-248): fp32 master parameters - one flat tensor per group ([`special_key`] vectors,
  all other <=1-d parameters, all matrices; `get_param_groups_and_shapes` :87-110) - a dynamic loss scale 2**lg_loss_scale
  (`backward` scales the loss, `optimize` skips the step and lowers the exponent by one on overflow, otherwise unscales, steps,
  copies the masters back into the fp16 model parameters and grows the exponent by `fp16_scale_growth`).

On this path the model parameters are the drop-in `UNetModel`'s (fp16 torso convolutions after `convert_to_fp16()`); the copy
back bumps their version counters, so the next forward re-packs the bf16 tensor-core operands (native.NativeNet).  The
gradient-norm reduction and the unscale run on the device (two reductions + one multiply per group instead of a host
round trip per tensor); the overflow decision needs ONE host read per step, like the reference's `.item()` chain.
"""
import math

import torch
import torch.nn as nn

INITIAL_LOG_LOSS_SCALE = 20.0


def convert_module_to_f16(l):
    if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
        l.weight.data = l.weight.data.half()
        if l.bias is not None:
            l.bias.data = l.bias.data.half()


def convert_module_to_f32(l):
    if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
        l.weight.data = l.weight.data.float()
        if l.bias is not None:
            l.bias.data = l.bias.data.float()


def get_param_groups_and_shapes(named_model_params, special_key=None):
    """[(named params, master shape)]: optionally the `special_key` vectors first, then every parameter with ndim <= 1 as one
    flat vector, then every matrix / kernel as one [1, -1] row (fp16_util.py:87-110)."""
    named = list(named_model_params)
    groups = []
    if special_key is not None:
        groups.append(([(n, p) for n, p in named if special_key in n], (-1)))
        named = [(n, p) for n, p in named if special_key not in n]
    groups.append(([(n, p) for n, p in named if p.ndim <= 1], (-1)))
    groups.append(([(n, p) for n, p in named if p.ndim > 1], (1, -1)))
    return groups


def _flat(tensors):
    return torch.cat([t.reshape(-1) for t in tensors]) if tensors else torch.zeros(0)


def make_master_params(param_groups_and_shapes):
    masters = []
    for group, shape in param_groups_and_shapes:
        m = nn.Parameter(_flat([p.detach().float() for _, p in group]).view(shape))
        m.requires_grad = True
        masters.append(m)
    return masters


def unflatten_master_params(param_group, master_param):
    out, off = [], 0
    flat = master_param.view(-1)
    for _, p in param_group:
        n = p.numel()
        out.append(flat[off:off + n].view(p.shape))
        off += n
    return out


def model_grads_to_master_grads(param_groups_and_shapes, master_params):
    for m, (group, shape) in zip(master_params, param_groups_and_shapes):
        m.grad = _flat([(p.grad.detach() if p.grad is not None else torch.zeros_like(p)).float() for _, p in group]).view(shape)


def master_params_to_model_params(param_groups_and_shapes, master_params):
    for m, (group, _) in zip(master_params, param_groups_and_shapes):
        for (_, p), src in zip(group, unflatten_master_params(group, m.detach())):
            p.detach().copy_(src)  # in-place: bumps the version counter -> the B200 net re-packs its operand copies


def master_params_to_state_dict(model, param_groups_and_shapes, master_params, use_fp16):
    sd = model.state_dict()
    if use_fp16:
        for m, (group, _) in zip(master_params, param_groups_and_shapes):
            for (name, _), t in zip(group, unflatten_master_params(group, m.detach())):
                assert name in sd
                sd[name] = t
    else:
        for i, (name, _) in enumerate(model.named_parameters()):
            assert name in sd
            sd[name] = master_params[i]
    return sd


def state_dict_to_master_params(model, state_dict, use_fp16):
    if use_fp16:
        named = [(name, state_dict[name]) for name, _ in model.named_parameters()]
        return make_master_params(get_param_groups_and_shapes(named))
    return [state_dict[name] for name, _ in model.named_parameters()]


def zero_master_grads(master_params):
    for p in master_params:
        p.grad = None


def zero_grad(model_params):
    for p in model_params:
        if p.grad is not None:
            p.grad.detach_()
            p.grad.zero_()


def check_overflow(value):
    return value == float("inf") or value == -float("inf") or value != value


class MixedPrecisionTrainer:
    def __init__(self, *, model, use_fp16=False, fp16_scale_growth=1e-3, initial_lg_loss_scale=INITIAL_LOG_LOSS_SCALE,
                 special_key=None):
        self.model = model
        self.use_fp16 = use_fp16
        self.fp16_scale_growth = fp16_scale_growth
        self.model_params = list(model.parameters())
        self.master_params = self.model_params
        self.param_groups_and_shapes = None
        self.lg_loss_scale = initial_lg_loss_scale
        self.last_grad_norm = None
        self.last_param_norm = None
        if use_fp16:
            self.param_groups_and_shapes = get_param_groups_and_shapes(model.named_parameters(), special_key=special_key)
            self.master_params = make_master_params(self.param_groups_and_shapes)

    def zero_grad(self):
        zero_grad(self.model_params)

    def backward(self, loss):
        if self.use_fp16:
            (loss * 2 ** self.lg_loss_scale).backward()
        else:
            loss.backward()

    def optimize(self, opt):
        return self._optimize_fp16(opt) if self.use_fp16 else self._optimize_normal(opt)

    def _compute_norms(self, grad_scale=1.0):
        """(||grads|| / grad_scale, ||params||) over the master parameters: device reductions, one host read."""
        with torch.no_grad():
            dev = self.master_params[0].device
            acc = torch.zeros(2, device=dev, dtype=torch.float32)
            for p in self.master_params:
                acc[1] += p.detach().float().pow(2).sum()
                if p.grad is not None:
                    acc[0] += p.grad.detach().float().pow(2).sum()
            g2, p2 = acc.tolist()
        return math.sqrt(g2) / grad_scale if g2 == g2 else float("nan"), math.sqrt(p2)

    def _optimize_fp16(self, opt):
        model_grads_to_master_grads(self.param_groups_and_shapes, self.master_params)
        grad_norm, param_norm = self._compute_norms(grad_scale=2 ** self.lg_loss_scale)
        self.last_grad_norm, self.last_param_norm = grad_norm, param_norm
        if check_overflow(grad_norm):
            self.lg_loss_scale -= 1
            zero_master_grads(self.master_params)
            return False
        inv = 1.0 / (2 ** self.lg_loss_scale)
        for p in self.master_params:
            p.grad.mul_(inv)
        opt.step()
        zero_master_grads(self.master_params)
        master_params_to_model_params(self.param_groups_and_shapes, self.master_params)
        self.lg_loss_scale += self.fp16_scale_growth
        return True

    def _optimize_normal(self, opt):
        self.last_grad_norm, self.last_param_norm = self._compute_norms()
        opt.step()
        return True

    def master_params_to_state_dict(self, master_params):
        return master_params_to_state_dict(self.model, self.param_groups_and_shapes, master_params, self.use_fp16)

    def state_dict_to_master_params(self, state_dict):
        return state_dict_to_master_params(self.model, state_dict, self.use_fp16)
