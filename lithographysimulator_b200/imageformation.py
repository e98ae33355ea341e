"""Drop-in module name: ``from lithographysimulator_b200.imageformation import abbeImage`` (reference imageformation.py)."""
from .imaging import abbeImage, calculateAerial, calculateFFTAerial  # noqa: F401
from .optics import Mask  # noqa: F401  (the reference forgets this import)
