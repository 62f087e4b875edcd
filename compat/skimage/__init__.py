"""stand-in for scikit-image: io / color are imported by palette/utils.py for the palette-extraction step only"""
from . import io, color  # noqa: F401
