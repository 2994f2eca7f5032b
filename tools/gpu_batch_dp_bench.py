#!/usr/bin/env python3
"""BASELINE config 4: Hunyuan3D-2mini-Turbo decoder parameters (downsample 2, expand 1, no ln_post / q-k norm),
FlashVDMVolumeDecoding (adaptive KV selection, top-k 256 of 512 tokens) at octree 384, a batch of 8 latents data-parallel,
one whole mesh per GPU (hy3dgeo.parallel.latents2mesh_data_parallel; no data-path collective).  Low-frequency saturating
synthetic field.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node 8 tools/gpu_batch_dp_bench.py [--res 384] [--reps 3]
"""
import argparse, json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, _lib, parallel as P
from hy3dgeo.volume_decoders import VanillaVolumeDecoder, FlashVDMVolumeDecoding

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=384)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = W.MINI_TURBO
sd0 = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=1.0, bias=0.0)
vae = hy3dgeo.B200ShapeVAE(cfg, sd0, device=dev)
z = torch.cat([W.synthetic_latents(cfg, 1, 1234 + b) for b in range(args.batch)], 0).to(dev)      # seeds 1234 .. 1241
g0 = VanillaVolumeDecoder()(vae(z[:1]), vae.geo_decoder, bounds=1.01, octree_resolution=96)[0]
q85, q90 = [float(v) for v in torch.quantile(g0.flatten()[::7].float(), torch.tensor([0.85, 0.90], device=dev))]
gain = 1.9 / (q90 - q85); bias = -0.95 - gain * q85
sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=gain, bias=bias)
vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev, volume_decoder=FlashVDMVolumeDecoding("mean"))
kw = dict(bounds=1.01, mc_level=0.0, num_chunks=8000, octree_resolution=args.res, mc_algo="mc", enable_pbar=False)


def step():
    return P.latents2mesh_data_parallel(vae, vae(z), None, 0, **kw)


outs = step()
t = []
for _ in range(args.reps):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    outs = step()
    torch.cuda.synchronize(); dist.barrier(); t.append((time.perf_counter() - t0) * 1e3)
_lib.get_context(dev).check_watchdog()
if rank == 0:
    st = vae.volume_decoder.last_stats[0]
    print(json.dumps({"config": f"mini-turbo FlashVDM(mean) octree {args.res}, batch {args.batch}, {world} GPUs (whole meshes per GPU)",
                      "batch_latents2mesh_ms": round(float(np.median(t)), 2), "all": [round(x, 2) for x in t],
                      "meshes_per_s": round(args.batch / (float(np.median(t)) / 1e3), 1),
                      "levels": st["levels"], "queries_item0": st["queries"],
                      "meshes": [None if o is None else [int(o.mesh_v.shape[0]), int(o.mesh_f.shape[0])] for o in outs]}), flush=True)
dist.destroy_process_group()
