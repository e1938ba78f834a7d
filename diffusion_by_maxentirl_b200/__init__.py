"""B200-native (sm_100a) implementation of the DxMI few-step sampler rollout hot path.

Public surface mirrors the reference's module interfaces (see `models/` for the drop-in classes) on top of the
C-ABI library `libdxmi_b200.so` (`include/dxmi_b200.h`).  `install()` binds the drop-ins into an importable checkout
of the reference (swyoon/Diffusion-by-MaxEntIRL) so its unchanged scripts and YAML `_target_`s resolve to this path.
"""
import importlib

__version__ = "0.1.0"

# (reference module, attribute) -> (drop-in module, attribute): exactly the symbols on the hot path
_BINDINGS = [
    ("models.DxMI.unet_small", "Model", "models.DxMI.unet_small", "Model"),
    ("models.DxMI.var_sampler", "VARSampler", "models.DxMI.var_sampler", "VARSampler"),
    ("models.DxMI.openai_diffusion", "OpenAIDiffusion", "models.DxMI.openai_diffusion", "OpenAIDiffusion"),
    ("models.cm.unet", "UNetModel", "models.cm.unet", "UNetModel"),
    ("models.cm.karras_diffusion", "KarrasDenoiser", "models.cm.karras_diffusion", "KarrasDenoiser"),
    ("models.cm.script_util", "create_model", "models.cm.script_util", "create_model"),
    ("models.cm.script_util", "create_model_and_diffusion", "models.cm.script_util", "create_model_and_diffusion"),
    ("models.modules", "IGEBMEncoderV2", "models.modules", "IGEBMEncoderV2"),
    # callers either side of the path: the replay-buffer helpers of the trainer
    ("models.DxMI.trainer", "append_buffer", "models.DxMI.trainer", "append_buffer"),
    ("models.DxMI.trainer", "reset_buffer", "models.DxMI.trainer", "reset_buffer"),
]


def set_default_precision(mode):
    """See native.set_default_precision: "bf16" (default) or "fp32"."""
    from . import native

    native.set_default_precision(mode)


def install(trainer_ops=False):
    """Rebind the hot-path classes of the reference's `models` package (which must be importable, i.e. the reference
    checkout is on sys.path) to the B200 drop-ins, in place.  Everything else in the reference (trainer, CLIs, FID,
    loaders, `models.value.TimeIndependentValue`, ...) is left untouched and keeps calling the same names.
    trainer_ops=True additionally binds the fused trainer-side ops of SURVEY 8f rank 1 (`train_ops.py`):
    `DxMI_Trainer{,_Cond}.get_running_cost` -> the fused forward/backward kernel, `models.cm.fp16_util.MixedPrecisionTrainer`
    -> the drop-in with device-side norm reductions (the reference's own class also works unchanged on the drop-in U-Net), and
    `torch.optim.Adam` AS SEEN BY the
    reference's train scripts is left alone - pass `diffusion_by_maxentirl_b200.train_ops.FusedAdam` explicitly where the
    script builds its optimizers (train_cifar10.py:283-296) to get the multi-tensor clip + Adam step.
    Returns the list of rebound `module.attr` names."""
    done = []
    for ref_mod, ref_attr, our_mod, our_attr in _BINDINGS:
        ref = importlib.import_module(ref_mod)
        ours = importlib.import_module(__name__ + "." + our_mod)
        setattr(ref, ref_attr, getattr(ours, our_attr))
        done.append(f"{ref_mod}.{ref_attr}")
    if trainer_ops:
        from . import train_ops

        tr = importlib.import_module("models.DxMI.trainer")
        for cls in ("DxMI_Trainer", "DxMI_Trainer_Cond"):
            if hasattr(tr, cls):
                setattr(getattr(tr, cls), "get_running_cost", train_ops.trainer_get_running_cost)
                done.append(f"models.DxMI.trainer.{cls}.get_running_cost")
        # EDM training: the same master-parameter / loss-scale contract with device-side norm reductions
        from .models.cm import fp16_util as ours_fp16

        ref_fp16 = importlib.import_module("models.cm.fp16_util")
        ref_fp16.MixedPrecisionTrainer = ours_fp16.MixedPrecisionTrainer
        done.append("models.cm.fp16_util.MixedPrecisionTrainer")
    return done
