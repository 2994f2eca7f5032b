"""CPU, world_size 2, gloo: host-side sharding logic of hy3dgeo.parallel (slab partition, ordered
list split, gathers) with an analytic field standing in for the decoder kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hy3dgeo  # noqa: F401
from hy3dgeo import parallel as P
from oracle import mc as OM, volume as OV


def field(p):
    return torch.tanh(20 * (0.6 - p.float().norm(dim=-1)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = 20
        N = res + 1
        pts = torch.from_numpy(OV.dense_points(1.01, res))

        def decode_range(first, count, out):
            out[:count] = field(pts[first:first + count])
        full = P.decode_dense_sharded(decode_range, (N, N, N), "cpu", None, to_all=True)
        root = P.decode_dense_sharded(decode_range, (N, N, N), "cpu", None, to_all=False)
        ref = OV.vanilla_decode(field, 1.01, 10 ** 6, res)
        ok = np.array_equal(full.numpy(), ref) and ((root is None) if rank else np.array_equal(root.numpy(), ref))
        # refined level: ordered list split across ranks
        up = OV.refine_active_set(ref, 0.0, True)
        idx = torch.from_numpy(np.flatnonzero(up.reshape(-1)).astype(np.int32))
        n = 2 * N - 1

        def values(sl):
            ijk = np.stack(np.unravel_index(sl.numpy(), (n, n, n)), 1)
            return field(torch.from_numpy(OV.refined_coords(ijk, 1.01, 2 * res)))
        vals = P.decode_list_sharded(values, idx, None)
        ijk = np.stack(np.unravel_index(idx.numpy(), (n, n, n)), 1)
        ok = ok and np.array_equal(vals.numpy(), field(torch.from_numpy(OV.refined_coords(ijk, 1.01, 2 * res))).numpy())
        class FakeVAE:                                   # whole meshes per rank (config 4)
            def latents2mesh(self, lat, **kw):
                return [("mesh", int(lat[0, 0, 0]), kw["octree_resolution"])]
        outs = P.latents2mesh_data_parallel(FakeVAE(), torch.arange(5.).view(5, 1, 1), None, 0, octree_resolution=9)
        ok = ok and (outs == [("mesh", b, 9) for b in range(5)] if rank == 0 else outs is None)
        # sharded marching cubes (config 5): halo exchange, global vertex ids, mesh gather — oracle MC as the slab extractor
        x0, x1 = P.slab_planes(N, rank, world)
        full_np = ref.astype(np.float32)
        state = {}

        def count_slab(slab, own):
            sub = slab.numpy()
            try:
                v, f, _, _ = OM.marching_cubes(sub, 0.0)
            except (ValueError, RuntimeError):
                state.update(v=np.zeros((0, 3), np.float32), f=np.zeros((0, 3), np.int32))
                return 0, 0, (float(sub.min()), float(sub.max()), False)
            owned = int((np.floor(v[:, 0]) < own).sum())              # vertices are in lexicographic voxel order: a prefix
            if own + 1 < sub.shape[0]:                                # triangles of the cubes based at owned planes
                try:
                    vt, ft, _, _ = OM.marching_cubes(sub[:own + 1], 0.0)
                    lut = {tuple(p): i for i, p in enumerate(map(tuple, v))}
                    ft = np.array([[lut[tuple(vt[i])] for i in tri] for tri in ft], np.int32).reshape(-1, 3)
                except (ValueError, RuntimeError):
                    ft = np.zeros((0, 3), np.int32)
            else:
                ft = f
            state.update(v=v[:owned], f=ft)
            return owned, len(ft), (float(sub.min()), float(sub.max()), False)

        def emit_slab(nv, nf, plane0, id_base):
            v = state["v"].copy(); v[:, 0] += plane0
            return torch.from_numpy(v), torch.from_numpy(state["f"] + np.int32(id_base))
        mesh = P.extract_mesh_sharded(torch.from_numpy(full_np[x0:x1].copy()), x0, count_slab, emit_slab, 0.0, None, 0)
        if rank == 0:
            v_ref, f_ref, _, _ = OM.marching_cubes(full_np, 0.0)
            # faces exact; vertices to 1 ulp (this CPU stand-in adds the plane offset after the float32 rounding — the CUDA
            # kernels interpolate in the global index frame and are bit-exact, tests/test_gpu_parity.py)
            ok = ok and np.array_equal(mesh[1].numpy(), f_ref) and mesh[0].shape == v_ref.shape and \
                bool(np.abs(mesh[0].numpy() - v_ref).max() < 2e-6)
        else:
            ok = ok and mesh is None
        lat = P.broadcast_latents(torch.arange(6.).view(2, 3) if rank == 0 else None, (2, 3), "cpu")
        ok = ok and bool((lat == torch.arange(6.).view(2, 3)).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert results == {0: True, 1: True}


def test_partitions_cover_exactly():
    for n in (1, 7, 129, 257, 385):
        for world in (1, 2, 3, 4, 8):
            spans = [P.slab_planes(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
