"""hy3dgeo — B200-native (sm_100a) geometry decoding for Hunyuan3D-2.

Drop-in replacements for the reference's ``volume_decoder`` / ``surface_extractor``
plugin slots (reference hy3dgen/shapegen/models/autoencoders/model.py:92-110),
backed by hand-written CUDA kernels behind the C-ABI declared in
``include/hy3dgeo.h``.  There is no CPU fallback: every compute entry point
raises if ``libhy3dgeo.so`` is missing.
"""
from . import weights  # noqa: F401

__all__ = ["weights"]
