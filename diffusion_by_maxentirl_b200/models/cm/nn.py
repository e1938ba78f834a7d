"""Drop-in for the hot subset of the reference's models/cm/nn.py: `append_dims` (:95-102).  GroupNorm32, SiLU and
`timestep_embedding` live inside the CUDA plan (csrc/kernels.cu)."""


def append_dims(x, target_dims):
    """Append trailing singleton dimensions until `x` has `target_dims` dimensions."""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]
