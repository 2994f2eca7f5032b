#!/usr/bin/env python3
"""Parity at scale: the tcgen05 decoder (product path) against the fp32 CUDA-core statement of the same decoder
(HY3D_PRECISION_FP32_SIMT, itself within 1.4e-6 of the CPU oracle in the tests) on 2^20 random query points of the full
model, plus the agreement of the occupancy sign (what marching cubes classifies) on those points.  One JSON line."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, _lib
from hy3dgeo.volume_decoders import bind

dev = torch.device("cuda:0")
out = {}
for name, cfg in (("full", W.FULL), ("mini", W.MINI), ("mini_turbo", W.MINI_TURBO)):
    sd = W.synthetic_state_dict(cfg, seed=0)
    vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
    lat = vae(W.synthetic_latents(cfg, 1, 1234).to(dev))
    ctx = bind(lat, vae.geo_decoder)
    n = 1 << 20
    xyz = (torch.rand(n, 3, generator=torch.Generator().manual_seed(11)) * 2.02 - 1.01).to(dev)
    ctx.set_precision(_lib.PRECISION_FP32_SIMT); ctx.prepare_kv(lat[0]); ref = ctx.decode_points(xyz).double()
    ctx.set_precision(_lib.PRECISION_FP16_TC); ctx.prepare_kv(lat[0]); tc = ctx.decode_points(xyz).double()
    ctx.check_watchdog()
    d = (tc - ref).abs()
    out[name] = {"points": n, "logit_range": [float(ref.min()), float(ref.max())], "max_abs_err": float(d.max()), "rms_err": float(d.pow(2).mean().sqrt()),
                 "frac_above_1e-3": float((d > 1e-3).double().mean()), "sign_disagreements": int(((tc > 0) != (ref > 0)).sum()),
                 "min_abs_logit_where_sign_differs": float(ref.abs()[(tc > 0) != (ref > 0)].min()) if bool(((tc > 0) != (ref > 0)).any()) else None}
print(json.dumps(out))
