"""rbr_b200 — B200-native (sm_100a) review-encoder hot path for DeepCoNN / NARRE / D-ATT.

Host side: PyTorch modules mirroring the reference's nn.Module surface
(models/deepconn/deepconn.py, models/narre/narre.py, models/dual_att/dual_att.py of
H263/review-based-recommender).  Device side: hand-written CUDA kernels behind the C-ABI
declared in include/rbr_b200.h (csrc/librbr_b200.so), loaded with ctypes.  There is no CPU
fallback: calling a compute op without the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"

from . import synth  # noqa: F401  (pure-torch, no native dependency)


def __getattr__(name):          # lazy: model classes import torch.nn and the native binding on first use
    if name == "DeepCoNNpp":
        from .deepconn import DeepCoNNpp
        return DeepCoNNpp
    if name == "NARRE":
        from .narre import NARRE
        return NARRE
    if name == "SimpleSiamese":
        from .simple_siamese import SimpleSiamese
        return SimpleSiamese
    if name == "DualAtt":
        from .dual_att import DualAtt
        return DualAtt
    if name in ("layers", "ops", "parallel", "graphs", "deepconn", "narre", "dual_att", "_lib", "staging", "optim", "inference", "simple_siamese"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
