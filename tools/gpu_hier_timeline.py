#!/usr/bin/env python3
"""Per-rank stage timeline of BASELINE configs[2] (Hierarchical octree 384, sharded) under torchrun: where the time that
does not shrink with the number of GPUs goes.  Stages are separated by device synchronisations (so the sum is a little
above the free-running step); every rank's table is gathered and printed by rank 0 as one JSON line.

    torchrun --nproc-per-node N tools/gpu_hier_timeline.py [--steps 5]
"""
import argparse, json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W, parallel as P
import bench as B

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--res", type=int, default=384)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = W.FULL
sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0), cfg, **B.SPARSE_FULL)
vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
dec = P.ShardedHierarchicalVolumeDecoding(timeline=True)
vae.volume_decoder = dec
z = W.synthetic_latents(cfg, 1, 1234).to(dev)
kw = dict(bounds=1.01, mc_level=0.0, num_chunks=8000, octree_resolution=args.res, mc_algo="mc", enable_pbar=False)
for _ in range(3):
    vae.latents2mesh(vae(z, group=True), **kw)
dec.timeline.clear()
extra = {"latent transformer": 0.0, "marching cubes (slab) + mesh gather + D2H": 0.0}
for _ in range(args.steps):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    lat = vae(z, group=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    dec._t_last = None
    grid = dec(lat, vae.geo_decoder, **kw); torch.cuda.synchronize(); t2 = time.perf_counter()
    vae.surface_extractor(grid, **kw); torch.cuda.synchronize(); t3 = time.perf_counter()
    extra["latent transformer"] += (t1 - t0) * 1e3
    extra["marching cubes (slab) + mesh gather + D2H"] += (t3 - t2) * 1e3
tab = {k: round(v / args.steps, 3) for k, v in {**extra, **dec.timeline}.items()}
tab["rank_queries"] = dec.last_stats[0]["rank_queries"]
out = [None] * world
dist.all_gather_object(out, tab)
if rank == 0:
    print(json.dumps({"n_gpus": world, "queries": dec.last_stats[0]["queries"], "per_rank_ms": out}))
dist.destroy_process_group()
