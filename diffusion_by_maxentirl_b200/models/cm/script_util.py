"""Drop-in for the constructor contract of the reference's models/cm/script_util.py: `create_model_and_diffusion`
(:54-101) and `create_model` (:104-158)."""
from .karras_diffusion import KarrasDenoiser
from .unet import UNetModel

NUM_CLASSES = 1000


def create_model_and_diffusion(image_size, class_cond, learn_sigma, num_channels, num_res_blocks, channel_mult,
                               num_heads, num_head_channels, num_heads_upsample, attention_resolutions, dropout,
                               use_checkpoint, use_scale_shift_norm, resblock_updown, use_fp16,
                               use_new_attention_order, weight_schedule, sigma_min=0.002, sigma_max=80.0,
                               distillation=False):
    model = create_model(image_size, num_channels, num_res_blocks, channel_mult=channel_mult, learn_sigma=learn_sigma,
                         class_cond=class_cond, use_checkpoint=use_checkpoint,
                         attention_resolutions=attention_resolutions, num_heads=num_heads,
                         num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                         use_scale_shift_norm=use_scale_shift_norm, dropout=dropout, resblock_updown=resblock_updown,
                         use_fp16=use_fp16, use_new_attention_order=use_new_attention_order)
    diffusion = KarrasDenoiser(sigma_data=0.5, sigma_max=sigma_max, sigma_min=sigma_min, distillation=distillation,
                               weight_schedule=weight_schedule)
    return model, diffusion


def create_model(image_size, num_channels, num_res_blocks, channel_mult="", learn_sigma=False, class_cond=False,
                 use_checkpoint=False, attention_resolutions="16", num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0, resblock_updown=False, use_fp16=False,
                 use_new_attention_order=False):
    if channel_mult == "":
        if image_size == 512:
            channel_mult = (0.5, 1, 1, 2, 2, 4, 4)
        elif image_size == 256:
            channel_mult = (1, 1, 2, 2, 4, 4)
        elif image_size == 128:
            channel_mult = (1, 1, 2, 3, 4)
        elif image_size == 64:
            channel_mult = (1, 2, 3, 4)
        else:
            raise ValueError(f"unsupported image size: {image_size}")
    else:
        channel_mult = tuple(int(ch_mult) for ch_mult in channel_mult.split(","))
    attention_ds = [image_size // int(res) for res in attention_resolutions.split(",")]
    return UNetModel(image_size=image_size, in_channels=3, model_channels=num_channels,
                     out_channels=(3 if not learn_sigma else 6), num_res_blocks=num_res_blocks,
                     attention_resolutions=tuple(attention_ds), dropout=dropout, channel_mult=channel_mult,
                     num_classes=(NUM_CLASSES if class_cond else None), use_checkpoint=use_checkpoint,
                     use_fp16=use_fp16, num_heads=num_heads, num_head_channels=num_head_channels,
                     num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
                     resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order)
