#!/usr/bin/env python3
"""Latent-transformer error of the tcgen05 path vs the fp32 oracle for the attention kernel variants."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hy3dgeo
from hy3dgeo import weights as W, _lib
from oracle import decoder as OD
dev = torch.device('cuda:0')
for cfg, name in ((W.MINI, "mini"), (W.FULL, "full")):
    sd = W.synthetic_state_dict(cfg, seed=0)
    z = W.synthetic_latents(cfg, 1, 1234)
    torch.set_num_threads(os.cpu_count())
    ref = OD.shapevae_forward(sd, z, cfg.heads)
    vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
    ctx = _lib.get_context(dev)
    lt = vae(z.to(dev), impl="torch").cpu()
    print(name, "torch fp32", float((lt - ref).abs().max()), flush=True)
    for bits, poly in ((0, 1), (0, 0), (0x20, 0)):
        ctx.debug_experiment(bits, poly)
        lat = vae(z.to(dev)).cpu()
        d = (lat - ref).abs()
        print(name, hex(bits), poly, "max", float(d.max()), "rms", float(d.pow(2).mean().sqrt()), "|ref|max", float(ref.abs().max()), flush=True)
    ctx.debug_experiment(0, 2)
