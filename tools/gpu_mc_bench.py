#!/usr/bin/env python3
"""Marching-cubes / octree kernel timings at BASELINE grid sizes (CUDA events per kernel family,
many repetitions; inputs larger than L2 are cycled so that the field is read from HBM)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import _lib
dev = torch.device("cuda:0")
ctx = _lib.get_context(dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6551.0}
out = {}
for n in (257, 385, 513):
    x = torch.linspace(-1.01, 1.01, n, device=dev)
    r = torch.sqrt(x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2)
    nbuf = max(2, int(300e6 // (4 * n ** 3)) + 1)             # > 126 MB L2 in total
    grids = [(torch.tanh(20 * (0.6 - r)) + 0.001 * i).contiguous() for i in range(nbuf)]
    for g in grids[:2]:
        ctx.mc_count(g, 0.0)
    ctx.profile_read(); ctx.profile(True)
    reps = 20
    for i in range(reps):
        g = grids[i % nbuf]
        nv, nf, _ = ctx.mc_count(g, 0.0)
        v = torch.empty((nv, 3), dtype=torch.float32, device=dev); f = torch.empty((nf, 3), dtype=torch.int32, device=dev)
        ctx.mc_emit([n] * 3, [2.02] * 3, [-1.01] * 3, v, f)
    prof = ctx.profile_read(); ctx.profile(False)
    row = {k: prof[k][0] / prof[k][1] * 1e3 for k in ("mc_bits", "mc_rowcount", "mc_scan", "mc_emit")}   # us per launch
    gbs = 4.0 * n ** 3 / (row["mc_bits"] * 1e-6) / 1e9
    tot_us = sum(row.values())
    alg = 4.0 * n ** 3 + 12.0 * nv + 12.0 * nf
    out[n] = {"us": row, "mc_bits_GBps": gbs, "mc_bits_frac_hbm": gbs / peaks["hbm_gbs"], "V": nv, "F": nf,
              "mc_total_us": tot_us, "mc_total_GBps_algorithmic": alg / (tot_us * 1e-6) / 1e9}
    print(n, json.dumps(out[n]))
    # octree refine at the same fine size (coarse = (n+1)//2)
    nc = (n + 1) // 2
    coarse = grids[0][::2, ::2, ::2].contiguous()
    idx = torch.empty(nc ** 3 * 8 // 2, dtype=torch.int32, device=dev)
    ctx.refine_level(coarse, 0.0, True, idx)
    ctx.profile_read(); ctx.profile(True)
    for i in range(10):
        cnt = ctx.refine_level(coarse, 0.0, True, idx)
    prof = ctx.profile_read(); ctx.profile(False)
    us = prof["octree"][0] / 10 * 1e3
    alg = 4.0 * nc ** 3 + 4.0 * cnt                       # read coarse, write the index list
    print(n, f"refine {nc}^3 -> {n}^3: {us:.1f} us per level ({prof['octree'][1] // 10} launches), active {cnt}, "
          f"algorithmic {alg / 1e6:.1f} MB -> {alg / (us * 1e-6) / 1e9:.0f} GB/s")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mc_bench.json"), "w"), indent=1)
