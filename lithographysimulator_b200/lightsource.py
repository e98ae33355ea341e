"""Drop-in module name for the reference's lightsource.py."""
from .optics import LightSource  # noqa: F401
