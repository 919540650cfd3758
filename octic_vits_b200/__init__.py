"""octic_vits_b200 -- B200-native (sm_100a) octic ViT block hot path behind the reference's module API.

The compute path is liboctic_b200.so (hand-written CUDA, C ABI in include/octic_b200.h); this package is the thin
host-side mirror of the reference's operator interface.  No CPU fallback, no Triton, no multi-backend dispatch.
"""
__version__ = "0.1.0"
