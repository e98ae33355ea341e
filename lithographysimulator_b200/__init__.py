"""lithographysimulator_b200 -- B200-native partially coherent Abbe imaging.

Drop-in for the aerial-image path of quarterwave0/LithographySimulator: the same object API
(``Mask``, ``LightSource``, ``Pupil``) and entry points (``abbeImage``, ``calculateFFTAerial``,
``calculateAerial``) on hand-written sm_100a CUDA kernels behind the C ABI in
``include/litho_b200.h``.  CUDA only; no CPU fallback.
"""
from .imaging import AbbeEngine, abbeImage, calculateAerial, calculateFFTAerial, epsilon_n, source_shifts
from .optics import LightSource, Mask, Pupil

__all__ = ["Mask", "LightSource", "Pupil", "abbeImage", "calculateFFTAerial", "calculateAerial", "AbbeEngine",
           "epsilon_n", "source_shifts"]
__version__ = "0.1.0"
