#!/usr/bin/env python3
"""Wall-clock breakdown of one hierarchical octree-384 latents->mesh pass (host view, synchronised stages)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, _lib
from hy3dgeo.volume_decoders import (VanillaVolumeDecoder, axis_tables, bind, hierarchy_levels, normalize_bounds, refine_level, SENTINEL)
from hy3dgeo.surface_extractors import MCSurfaceExtractor
dev = torch.device("cuda:0")
cfg = W.FULL
sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=17.5, bias=3.5)
vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
z = W.synthetic_latents(cfg, 1, 1234).to(dev)
T = {}
def tick(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
for rep in range(3):
    lat = tick("transformer fp32 (torch)", lambda: vae(z))
    tick("transformer fp16 (torch)", lambda: vae(z, dtype=torch.float16))
    torch.backends.cuda.matmul.allow_tf32 = True
    tick("transformer tf32 (torch)", lambda: vae(z))
    torch.backends.cuda.matmul.allow_tf32 = False
    ctx = tick("bind", lambda: bind(lat, vae.geo_decoder))
    tick("prepare_kv", lambda: ctx.prepare_kv(lat[0]))
    levels = hierarchy_levels(384, 63); b6 = normalize_bounds(1.01); bmin, bsz = b6[:3], b6[3:] - b6[:3]
    n0 = levels[0] + 1
    grid = torch.empty((n0, n0, n0), device=dev)
    tick("level0 dense 97^3", lambda: ctx.decode_dense(axis_tables(1.01, levels[0]), 0, n0 ** 3, grid))
    for r in levels[1:]:
        index = tick(f"refine -> {r+1}^3", lambda: refine_level(ctx, grid, 0.0, r == levels[-1]))
        n = r + 1
        nxt = torch.empty((n, n, n), device=dev)
        tick(f"fill {n}^3", lambda: ctx.fill(nxt, SENTINEL))
        cell = (bsz / r).astype(np.float32)
        tick(f"decode list {index.numel()}", lambda: ctx.decode_list(index, index.numel(), (n, n, n), cell, bmin.astype(np.float32), nxt))
        grid = nxt
    tick("sentinel->nan", lambda: ctx.sentinel_to_nan(grid, SENTINEL))
    ext = MCSurfaceExtractor()
    v, f = tick("mc count+emit (device)", lambda: ext.run_device(grid, mc_level=0.0, bounds=1.01, octree_resolution=384))
    tick("mesh D2H + numpy", lambda: (v.cpu().numpy(), f.cpu().numpy()))
for k, v in T.items():
    print(f"{k:32s} {np.median(v):8.2f} ms   {[round(x, 2) for x in v]}")
