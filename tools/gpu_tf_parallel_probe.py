"""Sequence-parallel latent transformer under torchrun: device time, host launch time and summed kernel time per rank
(is the loop host-bound?).  python -m torch.distributed.run --nproc-per-node N tools/gpu_tf_parallel_probe.py"""
import json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hy3dgeo
from hy3dgeo import weights as W

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
dev = torch.device("cuda", lr)
cfg = W.FULL
sd = W.synthetic_state_dict(cfg, seed=0)
vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
z = W.synthetic_latents(cfg, 1, 1234).to(dev)
ctx = hy3dgeo._lib.get_context(dev)
for _ in range(3):
    vae(z, group=True)
torch.cuda.synchronize(); dist.barrier()
res = {}
for name, fn in (("parallel", lambda: vae(z, group=True)), ("single", lambda: vae(z))):
    fn(); torch.cuda.synchronize()
    ctx.profile_read(); ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    e0.record()
    for _ in range(5):
        t0 = time.perf_counter(); fn(); host += time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_read(); ctx.profile(False)
    res[name] = {"device_ms": round(e0.elapsed_time(e1) / 5, 3), "host_launch_ms": round(host / 5 * 1e3, 3),
                 "kernel_ms_sum": round(sum(ms for ms, c in prof.values()) / 5, 3),
                 "launches": int(sum(c for ms, c in prof.values()) / 5)}
    dist.barrier()
out = [None] * world
dist.all_gather_object(out, res)
if rank == 0:
    print(json.dumps({"world": world, "per_rank": out}))
dist.destroy_process_group()
