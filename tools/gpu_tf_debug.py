#!/usr/bin/env python3
"""tcgen05 latent transformer vs oracle / torch fp32 (run under gpurun)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, _lib
from hy3dgeo.volume_decoders import bind
from oracle import decoder as OD
dev = torch.device("cuda:0")
for tag, cfg in (("mini", W.MINI), ("full", W.FULL)):
    sd = W.synthetic_state_dict(cfg, seed=0)
    vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
    z = W.synthetic_latents(cfg, 1, 1234)
    lat_o = OD.shapevae_forward(sd, z, cfg.heads)
    lat_t = vae(z.to(dev), impl="torch")
    lat_c = vae(z.to(dev), impl="tc")
    ctx = _lib.get_context(dev)
    print(tag, "watchdog", ctx.watchdog()[:5], "nan", int(torch.isnan(lat_c).sum()))
    for name, l in (("torch fp32", lat_t), ("tcgen05", lat_c)):
        d = (l.cpu() - lat_o).abs()
        print(f"  {name:10s} max|d| {float(d.max()):.3e} rms {float(d.pow(2).mean().sqrt()):.3e}  (|lat| max {float(lat_o.abs().max()):.2f}, rms {float(lat_o.pow(2).mean().sqrt()):.3f})")
    for impl in ("torch", "tc"):
        for _ in range(2): vae(z.to(dev), impl=impl)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): vae(z.to(dev), impl=impl)
        torch.cuda.synchronize(); print(f"  {impl}: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
    # effect on decoder logits
    gsd, fr = W.geo_decoder_state(sd), W.fourier_frequencies(cfg)
    q = (torch.rand(1, 512, 3, generator=torch.Generator().manual_seed(7)) * 2 - 1) * 1.01
    ref = OD.geo_decoder_forward(gsd, q, lat_o, fr, cfg.dec_heads)[0, :, 0]
    for name, l in (("torch fp32 latents", lat_t), ("tcgen05 latents", lat_c)):
        c = bind(l, vae.geo_decoder); c.prepare_kv(l[0])
        out = c.decode_points(q[0].to(dev)).cpu()
        print(f"  decoder logits with {name}: max|d| vs oracle {float((out-ref).abs().max()):.3e} rms {float((out-ref).pow(2).mean().sqrt()):.3e}")
