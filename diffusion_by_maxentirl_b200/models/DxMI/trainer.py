"""Drop-ins for the replay-buffer helpers of the reference's models/DxMI/trainer.py (`append_buffer` :23-55, `reset_buffer`
:58-70): the same dict contents, dtypes and row order (step-major: all rows of step 0, then step 1, ...), built
with ONE concatenation per key instead of the reference's T `torch.cat` calls per key (O(T) instead of O(T^2) bytes moved).

The B200 rollout returns `l_sample` / `mean` / `control` as views of one [T(+1), B, C, H, W] tensor each, so the per-key stack below
is usually a reshape of memory that already exists.  Everything else in the reference's trainer keeps running the reference's code.
"""
import torch


def reset_buffer(device):
    """Empty buffer dict with the reference's keys and dtypes (trainer.py:58-70)."""
    f = lambda: torch.empty(0, dtype=torch.float32, device=device)  # noqa: E731
    i = lambda: torch.empty(0, dtype=torch.long, device=device)  # noqa: E731
    return {"state": f(), "next_state": f(), "timestep": i(), "final": f(), "logp": f(), "control": f(), "entropy": f(),
            "mean": f(), "sigma": f(), "y": i()}


def _stack_steps(seq, n_seq):
    """[x_0, ..., x_{n_seq-1}] -> one [n_seq * B, ...] tensor in step-major order (a view when the list is a split of one tensor)."""
    first = seq[0]
    base = getattr(first, "_base", None)
    if base is not None and base.dim() == first.dim() + 1 and base.shape[0] >= n_seq and all(
            s._base is base and s.data_ptr() == base[k].data_ptr() for k, s in enumerate(seq[:n_seq])):
        return base[:n_seq].reshape(-1, *first.shape[1:]).detach()
    return torch.cat([s.detach() for s in seq[:n_seq]])


def _append(buf, key, new):
    old = buf[key]
    # the first append copies too (the reference always goes through torch.cat): `new` may be a view of the rollout's own
    # tensors, which a GraphedRollout overwrites on its next replay
    buf[key] = new.clone() if old.numel() == 0 else torch.cat((old, new))


def append_buffer(state_buffer, d_sample):
    """Reference trainer.py:23-55: appends the T transitions of one rollout to the buffer."""
    x_seq = d_sample["l_sample"]
    n_sample = len(x_seq[0])
    n_seq = len(x_seq) - 1
    device = x_seq[0].device
    states = _stack_steps(x_seq, n_seq + 1)  # [(T+1) * B, ...]
    _append(state_buffer, "state", states[: n_seq * n_sample])
    _append(state_buffer, "next_state", states[n_sample:])
    _append(state_buffer, "timestep", torch.arange(n_seq, device=device).repeat_interleave(n_sample))
    _append(state_buffer, "final", x_seq[-1].detach().repeat(n_seq, *([1] * (x_seq[-1].dim() - 1))))
    for key in ("logp", "control", "entropy", "mean", "sigma"):
        if key in d_sample:
            _append(state_buffer, key, _stack_steps(d_sample[key], n_seq))
    if "y" in d_sample:
        _append(state_buffer, "y", d_sample["y"].detach().repeat(n_seq))
    return state_buffer
