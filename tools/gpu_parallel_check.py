#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/gpu_parallel_check.py : sharded decoders == single-GPU decoders."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hy3dgeo
from hy3dgeo import weights as W
from hy3dgeo.parallel import ShardedVanillaVolumeDecoder, ShardedHierarchicalVolumeDecoding

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = W.MINI
sd = W.sparsify_field(W.synthetic_state_dict(cfg, seed=0, with_transformer=False), cfg, 2, 6.0, -2.0)
gd = hy3dgeo.GeoDecoder({k: v.to(dev) for k, v in W.geo_decoder_state(sd).items()}, cfg)
lat = torch.randn(1, 512, 1024, generator=torch.Generator().manual_seed(3)).to(dev)
a = ShardedVanillaVolumeDecoder(keep_sharded=False)(lat, gd, bounds=1.01, octree_resolution=45)      # gathered: tensor on rank 0, None elsewhere
b = hy3dgeo.VanillaVolumeDecoder()(lat, gd, bounds=1.01, octree_resolution=45)
h = ShardedHierarchicalVolumeDecoding(keep_sharded=False)(lat, gd, bounds=1.01, octree_resolution=64, min_resolution=15)   # tensor on all ranks
hs = hy3dgeo.HierarchicalVolumeDecoding()(lat, gd, bounds=1.01, octree_resolution=64, min_resolution=15)
ok_v = (a is None) if rank else bool(torch.equal(a, b))
ok_h = bool(torch.equal(torch.isnan(h), torch.isnan(hs))) and float((torch.nan_to_num(h) - torch.nan_to_num(hs)).abs().max()) == 0.0
print(f"rank {rank}: vanilla sharded == single {ok_v}; hierarchical sharded == single {ok_h} (visited {int((~torch.isnan(h)).sum())})", flush=True)
dist.destroy_process_group()
