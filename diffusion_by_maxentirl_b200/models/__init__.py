"""Drop-in mirror of the reference's `models` package for the sampler-rollout path.

Same module paths, class names, constructor signatures, `forward` / `sample` / `sample_step` signatures, return
dict keys and `state_dict` layouts as swyoon/Diffusion-by-MaxEntIRL, so the Hydra `_target_` strings of the
reference's YAML configs (`models.DxMI.unet_small.Model`, `models.DxMI.var_sampler.VARSampler`,
`models.value.TimeIndependentValue`, `models.modules.IGEBMEncoderV2`) and `models.cm.script_util
.create_model_and_diffusion` resolve to the B200 path.  See INTEGRATION.md.
"""
