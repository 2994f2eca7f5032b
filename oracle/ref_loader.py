"""ORACLE (test infrastructure): import the *real* reference modules from
``/root/reference`` without its missing dependencies.

Works only where the reference tree is mounted (the build container).  Used by
``oracle/make_golden.py`` to generate ``tests/golden`` and by
``tests/test_reference_live.py`` (skipped when the tree is absent).  The
reference tree is never modified; nothing is copied from it.
"""
from __future__ import annotations

import importlib
import inspect
import os
import sys
import types

REF_ROOT = os.environ.get("HY3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "hy3dgen", "shapegen", "models", "autoencoders"))


def load(marching_cubes=None):
    """Returns a namespace with the reference modules ``ab`` (attention_blocks),
    ``ap`` (attention_processors), ``vd`` (volume_decoders), ``se``
    (surface_extractors), ``mo`` (model).  ``marching_cubes`` stands in for the
    absent ``skimage.measure.marching_cubes``."""
    R = os.path.join(REF_ROOT, "hy3dgen")
    for name, path in [("hy3dgen", R), ("hy3dgen.shapegen", R + "/shapegen"),
                       ("hy3dgen.shapegen.models", R + "/shapegen/models"),
                       ("hy3dgen.shapegen.models.autoencoders", R + "/shapegen/models/autoencoders")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        sk.measure = types.ModuleType("skimage.measure")
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = sk.measure
    if marching_cubes is None:
        from . import mc as _mc
        marching_cubes = _mc.marching_cubes
    sys.modules["skimage.measure"].marching_cubes = marching_cubes
    base = "hy3dgen.shapegen.models.autoencoders."
    ns = types.SimpleNamespace()
    ns.ab = importlib.import_module(base + "attention_blocks")
    ns.ap = importlib.import_module(base + "attention_processors")
    ns.vd = importlib.import_module(base + "volume_decoders")
    ns.se = importlib.import_module(base + "surface_extractors")
    ns.mo = importlib.import_module(base + "model")
    ns.PatchedHierarchicalVolumeDecoding = _patched_hierarchical(ns.vd)
    return ns


def _patched_hierarchical(vd):
    """The reference class with the two ``dtype=next_points.dtype`` on
    volume_decoders.py:263-264 replaced by ``dtype=torch.float32`` (the form
    FlashVDM uses at :395-396).  Unpatched, every refined query is (-1,-1,-1)
    (SURVEY §0.3)."""
    src = inspect.getsource(vd.HierarchicalVolumeDecoding)
    assert src.count("dtype=next_points.dtype") == 2
    src = src.replace("dtype=next_points.dtype", "dtype=torch.float32")
    src = src.replace("class HierarchicalVolumeDecoding", "class PatchedHierarchicalVolumeDecoding")
    scope = dict(vd.__dict__)
    exec(compile(src, "<patched volume_decoders.py:185-277>", "exec"), scope)
    return scope["PatchedHierarchicalVolumeDecoding"]


def build_shapevae(ns, cfg, state_dict):
    """Reference ``ShapeVAE(**cfg)`` in eval/fp32 with ``state_dict`` loaded strictly."""
    import torch
    vae = ns.mo.ShapeVAE(**cfg.as_kwargs())
    vae.load_state_dict(state_dict, strict=True)
    return vae.eval().to(torch.float32)
