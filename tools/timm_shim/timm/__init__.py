"""Minimal stand-in for timm==1.0.12 (absent in this image, no network).

Used ONLY by tools/make_golden.py, in the build container, to import the reference package from /root/reference and
dump golden vectors.  It provides just the symbols the reference's octic path touches."""
from . import layers, models  # noqa: F401
from .models import create_model  # noqa: F401
