#!/usr/bin/env python3
"""BASELINE config 5: dense VanillaVolumeDecoder at octree 512 (or --res) sharded in axis-0 slabs, marching cubes per slab
behind a two-plane halo exchange, mesh pieces gathered on rank 0 (hy3dgeo.parallel.vanilla_latents2mesh_sharded), next
to the gather-the-grid variant (ShardedVanillaVolumeDecoder + MC on rank 0).  A low-frequency saturating synthetic field
(SURVEY §8d recipe) keeps the mesh at a realistic size.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node N tools/gpu_dense_sharded_bench.py [--res 512] [--reps 3]
"""
import argparse, json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, _lib, parallel as P
from hy3dgeo.volume_decoders import VanillaVolumeDecoder

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true", help="compare the sharded mesh with single-GPU MC of the gathered grid (bit-exact)")
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = W.FULL
sd0 = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=1.0, bias=0.0)
vae = hy3dgeo.B200ShapeVAE(cfg, sd0, device=dev)
z = W.synthetic_latents(cfg, 1, 1234).to(dev)
g0 = VanillaVolumeDecoder()(vae(z), vae.geo_decoder, bounds=1.01, octree_resolution=96)[0]
q85, q90 = [float(v) for v in torch.quantile(g0.flatten()[::7].float(), torch.tensor([0.85, 0.90], device=dev))]
gain = 1.9 / (q90 - q85); bias = -0.95 - gain * q85
sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, keep_freqs=2, gain=gain, bias=bias)
vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
kw = dict(bounds=1.01, mc_level=0.0, octree_resolution=args.res)


def slab_path():
    return P.vanilla_latents2mesh_sharded(vae(z), vae.geo_decoder, vae.surface_extractor, None, **kw)


def gather_path():
    grid = P.ShardedVanillaVolumeDecoder(keep_sharded=False)(vae(z), vae.geo_decoder, **kw)     # gathered on rank 0
    return vae.surface_extractor(grid, **kw) if rank == 0 else None


res = {"config": f"dense Vanilla octree {args.res}, full model, {world} GPUs", "queries": (args.res + 1) ** 3}
for name, fn in (("slab_mc", slab_path), ("gather_grid", gather_path)):
    outs = fn()
    t = []
    for _ in range(args.reps):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        outs = fn()
        torch.cuda.synchronize(); dist.barrier(); t.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        res[name] = {"latents2mesh_ms": round(float(np.median(t)), 2), "all": [round(x, 2) for x in t],
                     "mesh": [int(outs[0].mesh_v.shape[0]), int(outs[0].mesh_f.shape[0])] if outs[0] is not None else None}
        res[name + "_mesh"] = outs[0]
if rank == 0:
    a, b = res.pop("slab_mc_mesh"), res.pop("gather_grid_mesh")
    if args.check:
        res["slab_equals_whole"] = bool(a is not None and b is not None and np.array_equal(a.mesh_v.view(np.uint32), b.mesh_v.view(np.uint32))
                                        and np.array_equal(a.mesh_f, b.mesh_f))
    print(json.dumps(res), flush=True)
_lib.get_context(dev).check_watchdog()
dist.destroy_process_group()
