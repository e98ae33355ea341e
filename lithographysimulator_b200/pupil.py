"""Drop-in module name for the reference's pupil.py."""
from .optics import OSA, OSAindexToMN, Pupil, generatePhi, generateWavefrontError, generateZ  # noqa: F401
