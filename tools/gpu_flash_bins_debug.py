#!/usr/bin/env python3
"""Diagnose FlashVDM (mean) differences per spatial bin at the last level: device vs the oracle on the SAME latents.
Prints, per bin with a logit above tolerance: query count, samples, max error, selected-token set differences per head,
and the similarity gap at the top-k boundary (how close the tie is)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, _lib
from hy3dgeo.volume_decoders import FlashVDMVolumeDecoding
from oracle import decoder as OD, volume as OV

tag, res, mode = (sys.argv + ["turbo", "64", "mean"])[1:4]
res = int(res)
cfg = {"turbo": W.MINI_TURBO, "mini": W.MINI, "full": W.FULL}[tag]
kf, gain, bias = {"turbo": (2, 60.0, -16.0), "mini": (4, 4.0, 0.3), "full": (2, 6.0, 1.5)}[tag]
dev = torch.device("cuda:0")
sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, kf, gain, bias)
vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
z = W.synthetic_latents(cfg, 1, 1234).to(dev)
lat = vae(z, impl="torch")
ctx = _lib.get_context(dev)
dec = FlashVDMVolumeDecoding(mode, keep_levels=True)
out = dec(lat, vae.geo_decoder, bounds=1.01, num_chunks=600, mc_level=0.0, octree_resolution=res, min_resolution=15, enable_pbar=False)[0].cpu().numpy()
T = 256 if cfg.num_latents == 512 else 1024
H = cfg.dec_heads
sel_dev = ctx.flash_selection(216 * H * T).cpu().numpy().reshape(216, H, T)           # last level's selection
gsd, fr = W.geo_decoder_state(sd), W.fourier_frequencies(cfg)
lat_c = lat.cpu()
proc = OD.FlashProcessorOracle(mode)
calls = []

def dec_group(p, topk):
    proc.topk = topk
    o = OD.geo_decoder_forward(gsd, p, lat_c.expand(p.shape[0], -1, -1), fr, cfg.dec_heads, kv_select=proc)[..., 0]
    if topk is not True:
        calls.append((list(topk[0]), list(topk[1]), [s[0].numpy() for s in proc.last_selection], p[0].numpy().copy()))
    return o
torch.set_num_threads(os.cpu_count())
ref, st = OV.flashvdm_decode(dec_group, 1.01, 600, 0.0, res, 15, return_stats=True)
levels = st["levels"]
n = levels[-1] + 1
# last level bins of the oracle
nlast = sum(1 for c in st["calls"][-1])
last_calls = calls[-nlast:]
both = ~np.isnan(out) & ~np.isnan(ref)
err = np.where(both, np.abs(out - ref), 0.0)
print("max err", err.max(), "frac > tol", (err[both] > 1e-3 * gain).mean(), "visited equal", np.array_equal(np.isnan(out), np.isnan(ref)))
# k (fp32) and sampled q for the tie gap
k, v = OD.kv_heads(gsd, lat_c, H)
rows = []
for ids, cnts, sels, pts in last_calls:
    start = 0
    for b, c, s in zip(ids, cnts, sels):
        P = pts[start:start + c]
        ijk = np.rint((P - (-1.01)) / (2.02 / levels[-1])).astype(int)
        e = err[ijk[:, 0], ijk[:, 1], ijk[:, 2]]
        nd = [len(set(s[h]) ^ set(sel_dev[b, h])) // 2 for h in range(H)]
        if e.max() > 1e-3 * gain or sum(nd):
            x0 = OD._lin(OD.fourier_embed(torch.from_numpy(P[None, ::50]), fr), gsd, "query_proj")
            q = OD.q_heads(gsd, x0, H)
            sim = (q @ k.transpose(-1, -2)).mean(-2)[0]
            srt = torch.sort(sim, dim=-1, descending=True).values
            gap = ((srt[:, T - 1] - srt[:, T]) / (srt[:, 0] - srt[:, -1])).numpy()
            rows.append({"bin": int(b), "count": int(c), "samples": int((c + 49) // 50), "max_err": float(e.max()), "tokens_differ_per_head": nd,
                         "rel_gap_at_boundary": [float(f"{g:.2e}") for g in gap]})
        start += c
print(json.dumps(rows[:40], indent=0))
print("bins with differences:", len(rows))
