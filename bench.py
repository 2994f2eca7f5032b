#!/usr/bin/env python3
"""Benchmark of the geometry-decoding hot path (latents -> occupancy grid -> mesh).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one latents->mesh pass of BASELINE.json configs[1]: Hunyuan3D-2 ShapeVAE (3072 latent
tokens, width 1024, random-init weights, synthetic latent), VanillaVolumeDecoder at octree
resolution 256 (257^3 = 16 974 593 decoder queries) followed by marching cubes.  Metric: decoder
query points per second over the whole job (BASELINE.json: "latents->mesh ms and query pts/sec");
``ms_per_step`` is the latents->mesh time.  Prints ONE JSON line on rank 0.

``value``  : inputs resident in HBM, mesh left on the device (device-timed, CUDA events).
``e2e``    : the public API with HOST buffers: latents from pinned host memory, the mesh returned as
             numpy arrays (``B200ShapeVAE.latents2mesh``), copies inside the timed region.
``roofline``: the kernel family with the largest share of the step, algorithmic FLOPs / CUDA-event time.
``cpu_baseline`` / ``--impl reference``: the oracle port of the reference PyTorch path on the host cores,
             bounded sample (the reference tree itself is absent on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoder_query_points_per_sec"
UNIT = "pts/s"
CHUNK = 8000          # the reference's num_chunks default in pipelines (pipelines.py:689-693)


def flops_per_point(W, M, r, E=51):
    """F_pt = 2*51*W + 4*W^2 + 4*M*W + 4*r*W^2 + 2*W."""
    return {"gemm_query_proj": 2 * E * W, "gemm_c_q": 2 * W * W, "attention": 4 * M * W, "gemm_c_proj": 2 * W * W,
            "gemm_c_fc": 2 * r * W * W, "gemm_mlp_proj": 2 * r * W * W, "head": 2 * W}


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(self.rows), "reasons": reasons}


def cpu_port_setup(cfg, sd, z):
    """Oracle latent transformer on the host (once per latent): returns (latents, seconds)."""
    from oracle import decoder as OD
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    lat = OD.shapevae_forward(sd, z, cfg.heads)
    return lat, time.time() - t0


def cpu_port_rate(cfg, sd, lat, seconds, min_chunks=2, res=256):
    """Oracle port of the reference path on the host cores: chunks of 8000 dense-grid queries of the
    same workload through CrossAttentionDecoder.forward semantics (K/V re-projected per chunk, as
    the reference does, attention_blocks.py:251-257).  Returns (pts/s, chunks, threads)."""
    from hy3dgeo import weights as W
    from oracle import decoder as OD, volume as OV
    torch.set_num_threads(os.cpu_count())
    gsd, fr = W.geo_decoder_state(sd), W.fourier_frequencies(cfg)
    N = res + 1
    ax = OV.axis_tables(1.01, res)
    n, t0, done = 0, time.time(), 0
    while done < min_chunks or time.time() - t0 < seconds:
        lin = (np.arange(done * CHUNK, (done + 1) * CHUNK) + N * N * (N // 3)) % (N ** 3)     # interior planes of the grid
        k = lin % N; j = (lin // N) % N; i = lin // (N * N)
        pts = torch.from_numpy(np.stack([ax[0][i], ax[1][j], ax[2][k]], 1))
        with torch.no_grad():
            OD.geo_decoder_forward(gsd, pts[None], lat, fr, cfg.dec_heads)
        done += 1
        n += CHUNK
    dt = time.time() - t0
    return n / dt, done, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=256, help="octree resolution (default: BASELINE configs[1])")
    ap.add_argument("--model", default="full", choices=["full", "mini"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import hy3dgeo
    from hy3dgeo import weights as W, _lib
    cfg = W.FULL if args.model == "full" else W.MINI
    N = args.res + 1
    npts = N ** 3
    workload = (f"Hunyuan3D-2{'mini' if args.model == 'mini' else ''} ShapeVAE ({cfg.num_latents} latent tokens, width {cfg.width}) "
                f"VanillaVolumeDecoder octree_resolution={args.res} + marching cubes")
    config = {"workload": workload, "queries_per_step": npts, "grid": [N, N, N], "bounds": 1.01, "mc_level": 0.0,
              "weights": "random-init seed 0 (hy3dgeo.weights.synthetic_state_dict)", "latent_seed": 1234,
              "partition": f"axis-0 slabs x{world}" if world > 1 else "single GPU",
              "l2": "per-step activations (GBs) and the grid exceed the 126 MB L2; no explicit flush"}

    # -------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        sd = W.synthetic_state_dict(cfg, seed=0)
        z = W.synthetic_latents(cfg, 1, 1234)
        chunks_per_step = 2
        lat, tf_s = cpu_port_setup(cfg, sd, z)                       # latent transformer: once per latent, reported separately
        rates = []
        for s in range(args.warmup + args.steps):
            t0 = time.time()
            r, done, threads = cpu_port_rate(cfg, sd, lat, 0.0, min_chunks=chunks_per_step, res=args.res)
            if s >= args.warmup:
                rates.append((done * CHUNK, time.time() - t0))
        pts = sum(a for a, _ in rates); secs = sum(b for _, b in rates)
        val = pts / secs
        sample = (f"{chunks_per_step} chunks of {CHUNK} dense-grid queries per step (of {npts}); oracle port of the reference PyTorch "
                  f"fp32 CPU path; latent transformer {tf_s:.1f} s once per latent, not in the rate")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "extrapolated_full_step_s": tf_s + npts / val}))
        return

    # -------------------------------------------------------------------------------- our arm
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from hy3dgeo.parallel import ShardedVanillaVolumeDecoder
    sd = W.synthetic_state_dict(cfg, seed=0)
    vae = hy3dgeo.B200ShapeVAE(cfg, sd, device=dev)
    if world > 1:
        vae.volume_decoder = ShardedVanillaVolumeDecoder()
    z_host = W.synthetic_latents(cfg, 1, 1234).pin_memory()
    z_dev = z_host.to(dev)
    ctx = _lib.get_context(dev)
    kw = dict(bounds=1.01, mc_level=0.0, num_chunks=CHUNK, octree_resolution=args.res, mc_algo="mc", enable_pbar=False)
    mesh_bytes = [0]

    def step_device():
        lat = vae(z_dev)
        grid = vae.volume_decoder(lat, vae.geo_decoder, **kw)
        if grid is not None:
            v, f = vae.surface_extractor.run_device(grid[0], mc_level=0.0, bounds=1.01, octree_resolution=args.res)
            return v.shape[0], f.shape[0]
        return 0, 0

    def step_e2e():
        lat = vae(z_host.to(dev, non_blocking=True))
        grid = vae.volume_decoder(lat, vae.geo_decoder, **kw)
        if grid is not None:
            outs = vae.surface_extractor(grid, **kw)
            mesh_bytes[0] = outs[0].mesh_v.nbytes + outs[0].mesh_f.nbytes
            return outs
        return None

    def timed(fn, steps, profile=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            ctx.profile_read(); ctx.profile(True)
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        prof = None
        if profile:
            prof = ctx.profile_read(); ctx.profile(False)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.launches - l0, prof

    for _ in range(max(args.warmup, 3)):
        nv, nf = step_device()
    ctx.check_watchdog()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, prof = timed(step_device, args.steps, profile=True)
    clocks = sampler.summary() if sampler else None
    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    ctx.check_watchdog()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = npts * args.steps / (ms / 1e3)
    fl = flops_per_point(cfg.dec_width, cfg.num_latents, cfg.geo_decoder_mlp_expand_ratio)
    pts_local = npts / world            # each rank times its own slab; rank 0's families are reported
    fams = {k: v for k, v in prof.items() if v[1] > 0}
    total_fam_ms = sum(v[0] for v in fams.values())
    top = max((k for k in fams if k in fl), key=lambda k: fams[k][0])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    top_ms, top_cnt = fams[top]
    achieved = fl[top] * pts_local * args.steps / (top_ms / 1e3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top)
    except Exception:
        pass
    all_tf = sum(fl.values()) * pts_local * args.steps / (sum(fams[k][0] for k in fams if k in fl or k in ("layernorm", "embed")) / 1e3) / 1e12
    roofline = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "peak_source": peak_src, "launches": top_cnt, "avg_launch_ms": top_ms / max(top_cnt, 1),
                "flops_per_point": fl[top], "share_of_step": top_ms / total_fam_ms,
                "decoder_chain_tflops": all_tf, "decoder_chain_frac": all_tf / peak_tf,
                "families_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in sorted(fams.items(), key=lambda kv: -kv[1][0])}}
    if "mc_bits" in fams and peaks.get("hbm_gbs"):
        b_ms, b_cnt = fams["mc_bits"]
        gbs = 4.0 * npts * b_cnt / (b_ms / 1e3) / 1e9
        roofline["mc_bits"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                               "avg_launch_ms": b_ms / b_cnt, "algorithmic_bytes": 4 * npts}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate", "data": "synthetic", "config": config,
            "latents2mesh_ms": ms / args.steps, "mesh": {"vertices": nv, "faces": nf},
            "roofline": roofline,
            "e2e": {"value": npts * args.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": z_host.numel() * z_host.element_size(), "d2h_bytes_per_step": mesh_bytes[0]},
            "gpu_launches": launches, "clocks": clocks}
    if world == 1:
        lat_cpu, tf_s = cpu_port_setup(cfg, sd, W.synthetic_latents(cfg, 1, 1234))
        rate, chunks, threads = cpu_port_rate(cfg, sd, lat_cpu, args.cpu_seconds, res=args.res)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{chunks} chunks of {CHUNK} queries of the same grid (of {npts}); oracle port of the reference "
                                          f"fp32 PyTorch CPU path; full step extrapolates to {tf_s + npts / rate:.0f} s (latent transformer {tf_s:.1f} s)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
