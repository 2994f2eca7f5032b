#!/usr/bin/env python3
"""Calibrate the synthetic sparse field of BASELINE configs 3/4: choose the output head's (gain, bias)
from level-0 quantiles so that the octree-384 Hierarchical decoder visits 10-14 % of the fine grid (the tanh-sphere
planning workload of SURVEY §8c visits 11-14 %).  Prints one JSON line per candidate; the chosen constants are
hard-coded in bench.py (the weights are seeded, so the numbers are reproducible).

    python tools/gpu_calibrate_sparse.py [--model full|turbo] [--res 384]
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W
from hy3dgeo.volume_decoders import VanillaVolumeDecoder, HierarchicalVolumeDecoding

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="full")
ap.add_argument("--res", type=int, default=384)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = {"full": W.FULL, "mini": W.MINI, "turbo": W.MINI_TURBO}[args.model]
sd0 = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=1.0, bias=0.0)
vae = hy3dgeo.B200ShapeVAE(cfg, sd0, device=dev)
z = W.synthetic_latents(cfg, 1, 1234).to(dev)
lat = vae(z)
g0 = VanillaVolumeDecoder()(lat, vae.geo_decoder, bounds=1.01, octree_resolution=96)[0].flatten().float()
qs = [0.70, 0.75, 0.78, 0.80, 0.82, 0.85, 0.90, 0.92, 0.95]
qv = {q: float(v) for q, v in zip(qs, torch.quantile(g0[::3], torch.tensor(qs, device=dev)))}
print(json.dumps({"model": args.model, "level0_quantiles": qv, "mean": float(g0.mean()), "std": float(g0.std())}), flush=True)
for qa, qb in [(0.85, 0.90), (0.80, 0.90), (0.78, 0.90), (0.75, 0.90), (0.80, 0.92), (0.75, 0.92), (0.70, 0.90)]:
    gain = 1.9 / (qv[qb] - qv[qa]); bias = -0.95 - gain * qv[qa]
    sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=gain, bias=bias)
    v2 = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
    dec = HierarchicalVolumeDecoding()
    v2.volume_decoder = dec
    outs = v2.latents2mesh(v2(z), bounds=1.01, mc_level=0.0, num_chunks=8000, octree_resolution=args.res, mc_algo="mc", enable_pbar=False)
    st = dec.last_stats[0]
    print(json.dumps({"qa": qa, "qb": qb, "gain": gain, "bias": bias, "queries": st["queries"], "total": int(sum(st["queries"])),
                      "visited_fraction_last": st["queries"][-1] / (st["levels"][-1] + 1) ** 3,
                      "logit_std": float(g0.std()) * gain,
                      "mesh": None if outs[0] is None else [int(outs[0].mesh_v.shape[0]), int(outs[0].mesh_f.shape[0])]}), flush=True)
