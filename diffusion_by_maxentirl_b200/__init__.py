"""B200-native (sm_100a) implementation of the DxMI few-step sampler rollout hot path.

Public surface mirrors the reference's module interfaces (see `models/` for the drop-in classes) on top of the
C-ABI library `libdxmi_b200.so` (`include/dxmi_b200.h`).
"""
__version__ = "0.1.0"
