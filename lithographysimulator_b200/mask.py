"""Drop-in module name for the reference's mask.py."""
from .optics import Mask  # noqa: F401
