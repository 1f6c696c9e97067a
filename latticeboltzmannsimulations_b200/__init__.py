"""B200-native D2Q9 lid-driven-cavity collide-and-stream step (drop-in for the hot path of
RaghuvirJonnagiri/LatticeBoltzmannSimulations).  Python host code over a C-ABI CUDA library; no CPU fallback."""
from ._capi import LBMError  # noqa: F401
from .solver import CavitySolver  # noqa: F401
from .cavity import run_cavity, datagen, save_dataset  # noqa: F401

__all__ = ["CavitySolver", "run_cavity", "datagen", "save_dataset", "LBMError"]
