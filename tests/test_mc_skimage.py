"""Marching cubes against REAL scikit-image output — runs only where `tests/golden/skimage_mc.npz` exists (made by
`tools/make_skimage_golden.py` on a machine that has scikit-image; this image has none, so here the tests skip and the
marching-cubes oracle stays "parity unpinned").  What is compared is what can be equal between the classic
256-case table used here and skimage's Lewiner variant: the vertex SET (one vertex per sign-change edge, same
interpolation), the surface up to the triangulation of ambiguous cubes (Chamfer, area), and — reported, and asserted only
on fields without ambiguous cubes — face count and Euler characteristic."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "skimage_mc.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/skimage_mc.npz absent: scikit-image unavailable offline")


def _fields():
    spec = importlib.util.spec_from_file_location("make_skimage_golden", os.path.join(ROOT, "tools", "make_skimage_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _check(name, v, f, g, m):
    vs = v[np.lexsort(v.T[::-1])]
    ref = g[name + "_verts_sorted"]
    assert vs.shape == ref.shape, (name, vs.shape, ref.shape)                       # one vertex per sign-change edge
    both = np.isfinite(vs).all(1) & np.isfinite(ref).all(1)
    assert np.array_equal(np.isfinite(vs), np.isfinite(ref))
    assert np.abs(vs[both] - ref[both]).max() < 1e-5                                # same interpolation on the same edges
    st = m.mesh_stats(v, f)
    assert abs(st["area"] - float(g[name + "_area"])) < 2e-2 * float(g[name + "_area"])
    if name in ("sphere49", "two_blobs49"):                                         # no ambiguous cubes: topology must agree
        assert st["F"] == int(g[name + "_F"]) and st["euler"] == int(g[name + "_euler"])
    return st, {k: int(g[f"{name}_{k}"]) for k in ("F", "euler")}


def test_oracle_mc_vs_skimage():
    from oracle import mc as OM
    m = _fields()
    g = np.load(GOLD)
    for name, (vol, level) in m.fields().items():
        v, f, _, _ = OM.marching_cubes(vol, level)
        print(name, _check(name, v, f, g, m))


@pytest.mark.gpu
def test_cuda_mc_vs_skimage():
    import torch
    from hy3dgeo import _lib
    m = _fields()
    g = np.load(GOLD)
    ctx = _lib.get_context(torch.device("cuda:0"))
    for name, (vol, level) in m.fields().items():
        t = torch.from_numpy(vol).cuda()
        nv, nf, _ = ctx.mc_count(t, level)
        v = torch.empty((nv, 3), dtype=torch.float32, device="cuda"); f = torch.empty((nf, 3), dtype=torch.int32, device="cuda")
        ctx.mc_emit((1, 1, 1), (1, 1, 1), (0, 0, 0), v, f)
        print(name, _check(name, v.cpu().numpy(), f.cpu().numpy(), g, m))
