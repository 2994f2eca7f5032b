"""Volume decoders — host-side mirror of the reference plugin slot 1
(``hy3dgen/shapegen/models/autoencoders/volume_decoders.py``), same class names,
call signatures and return conventions, backed by ``libhy3dgeo.so``.

    grid_logits = volume_decoder(latents, geo_decoder, bounds=, num_chunks=,
                                 octree_resolution=, mc_level=, enable_pbar=, **kw)

``geo_decoder`` is the live ``CrossAttentionDecoder`` (or ``hy3dgeo.model.GeoDecoder``);
it is *read* (state_dict + hyper-parameters), never called and never mutated.  An
arbitrary callable raises ``TypeError``: there is no CPU / eager fallback.

``num_chunks`` and ``enable_pbar`` are accepted and ignored by the dense and
hierarchical decoders (the device code tiles queries itself); FlashVDM uses
``num_chunks`` only where it changes the *result* (bin packing has no effect on
results, mini-grid batching has none either), i.e. nowhere.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

from . import weights as W
from ._lib import GeoContext, get_context

SENTINEL = -10000.0
NAN = float("nan")


def normalize_bounds(bounds) -> np.ndarray:
    """reference volume_decoders.py:158-161."""
    if isinstance(bounds, float):
        bounds = [-bounds, -bounds, -bounds, bounds, bounds, bounds]
    b = np.asarray(bounds, dtype=np.float64)
    if b.shape != (6,):
        raise ValueError("bounds must be a float or a 6-sequence [xmin,ymin,zmin,xmax,ymax,zmax]")
    return b


def axis_tables(bounds, res: int) -> List[np.ndarray]:
    """Per-axis coordinates of ``generate_dense_grid_points`` (reference :131-133):
    ``np.linspace(min, max, res+1, dtype=float32)``.  O(N) host work instead of the
    reference's O(N^3) meshgrid + H2D copy."""
    b = normalize_bounds(bounds)
    return [np.linspace(b[a], b[a + 3], int(res) + 1, dtype=np.float32) for a in range(3)]


def hierarchy_levels(octree_resolution: int, min_resolution: int = 63) -> List[int]:
    """reference :202-208."""
    res, r = [], int(octree_resolution)
    if r < min_resolution:
        res.append(r)
    while r >= min_resolution:
        res.append(r)
        r //= 2
    res.reverse()
    return res


def flash_levels(octree_resolution: int, min_resolution: int = 63, mini_grid_num: int = 4) -> List[int]:
    """reference :310-319."""
    res = hierarchy_levels(octree_resolution, min_resolution)
    res[0] = round(res[0] / mini_grid_num) * mini_grid_num - 1
    for i in range(1, len(res)):
        res[i] = res[0] * 2 ** i
    return res


def _decoder_key(geo_decoder):
    """(state_dict, cache key) of a CrossAttentionDecoder-like object — no device synchronisation: the key is the
    module's identity plus every tensor's storage address and in-place version counter.  (Edits made through
    ``param.data`` do not bump the counter: call ``hy3dgeo._lib.get_context(device).invalidate_weights()`` after such
    an edit.)"""
    if not (hasattr(geo_decoder, "state_dict") and hasattr(geo_decoder, "fourier_embedder")
            and hasattr(geo_decoder, "cross_attn_decoder")):
        raise TypeError(
            "geo_decoder must be a CrossAttentionDecoder (its weights are read and run by the CUDA kernels); "
            f"got {type(geo_decoder).__name__}. hy3dgeo has no CPU/eager fallback for arbitrary callables.")
    sd = geo_decoder.state_dict()
    key = (id(geo_decoder),) + tuple((k, v.data_ptr(), getattr(v, "_version", 0)) for k, v in sd.items())
    return sd, key


def bind(latents: torch.Tensor, geo_decoder) -> GeoContext:
    """Context of the latents' device with this decoder's weights resident.  The hyper-parameters are read (one small
    device->host copy of the Fourier frequencies) only when the weights are not already cached."""
    if not latents.is_cuda:
        raise RuntimeError("latents must live on a CUDA device (hy3dgeo has no CPU path)")
    ctx = get_context(latents.device)
    sd, key = _decoder_key(geo_decoder)
    if not ctx.has_decoder(key, geo_decoder):
        ctx.set_decoder(sd, W.config_from_geo_decoder(geo_decoder), key, owner=geo_decoder)
    return ctx


class VanillaVolumeDecoder:
    """Dense evaluation (reference volume_decoders.py:141-182) -> float32 [B,N,N,N],
    index [b, ix, iy, iz], z fastest."""

    @torch.no_grad()
    def __call__(self, latents: torch.Tensor, geo_decoder, bounds: Union[Tuple[float], List[float], float] = 1.01,
                 num_chunks: int = 10000, octree_resolution: int = None, enable_pbar: bool = True, **kwargs):
        ctx = bind(latents, geo_decoder)
        axes = axis_tables(bounds, octree_resolution)
        N = int(octree_resolution) + 1
        B = latents.shape[0]
        out = torch.empty((B, N, N, N), dtype=torch.float32, device=latents.device)
        for b in range(B):
            ctx.prepare_kv(latents[b])
            ctx.decode_dense(axes, 0, N * N * N, out[b])
        return out


def refine_level(ctx: GeoContext, grid: torch.Tensor, mc_level: float, last: bool, nf: int = None) -> torch.Tensor:
    """Ordered flat indices of the voxels of the fine grid [nf]^3 to query (reference :245-260).  ``nf`` is the next
    level's r + 1: 2n-1 when the resolution doubles exactly, 2n when the coarser level is an odd r // 2 (:202-208)."""
    n = grid.shape[0]
    nf = 2 * n - 1 if nf is None else int(nf)
    cap = min(nf ** 3, max(1 << 20, nf ** 3 // 3))
    index = torch.empty(cap, dtype=torch.int32, device=grid.device)
    cnt = ctx.refine_level(grid, mc_level, last, index, nf)
    if cnt > cap:
        index = torch.empty(cnt, dtype=torch.int32, device=grid.device)
        cnt = ctx.refine_level(grid, mc_level, last, index, nf)
    return index[:cnt]


class HierarchicalVolumeDecoding:
    """Coarse-to-fine decoding (reference volume_decoders.py:185-277) -> latents.dtype
    [B,N,N,N] with NaN at unvisited voxels.

    The reference builds refined coordinates in int64 (:262-264), which collapses
    every refined query to (-1,-1,-1); this implements the evident
    intent — the float32 form FlashVDM uses at :394-396 — and is validated against
    the reference class patched the same way.  The reference is batch-1 only
    (``squeeze(0)``); batches are looped here.

    ``keep_levels`` (diagnostics / parity tests): keep every level's grid of the last item (sentinel form, -10000 =
    unvisited) in ``last_levels``."""

    def __init__(self, keep_levels: bool = False):
        self.keep_levels = keep_levels
        self.last_levels = []

    @torch.no_grad()
    def __call__(self, latents: torch.Tensor, geo_decoder, bounds: Union[Tuple[float], List[float], float] = 1.01,
                 num_chunks: int = 10000, mc_level: float = 0.0, octree_resolution: int = None, min_resolution: int = 63,
                 enable_pbar: bool = True, **kwargs):
        ctx = bind(latents, geo_decoder)
        levels = hierarchy_levels(octree_resolution, min_resolution)
        b6 = normalize_bounds(bounds)
        bbox_min, bbox_size = b6[:3], b6[3:] - b6[:3]
        outs = []
        self.last_stats = []
        for b in range(latents.shape[0]):
            ctx.prepare_kv(latents[b])
            n0 = levels[0] + 1
            grid = torch.empty((n0, n0, n0), dtype=torch.float32, device=latents.device)
            ctx.decode_dense(axis_tables(bounds, levels[0]), 0, n0 ** 3, grid)
            queries = [n0 ** 3]
            kept = [grid]
            for r in levels[1:]:
                n = r + 1
                last = r == levels[-1]
                index = refine_level(ctx, grid, mc_level, last=last, nf=n)
                nxt = torch.empty((n, n, n), dtype=torch.float32, device=latents.device)
                # the last level is born with NaN at the unvisited voxels (reference :245-246 + :275 in one pass: the
                # 228 MB sentinel -> NaN sweep over a 385^3 grid is gone); intermediate levels keep the -10000 sentinel
                ctx.fill(nxt, NAN if last and not self.keep_levels else SENTINEL)
                cell = (bbox_size / r).astype(np.float32)           # reference :243,:394 float32(resolution)
                ctx.decode_list(index, index.numel(), (n, n, n), cell, bbox_min.astype(np.float32), nxt)
                grid = nxt
                queries.append(int(index.numel()))
                kept.append(grid)
            if self.keep_levels:
                self.last_levels = kept[:-1] + [kept[-1].clone()]
            if self.keep_levels or len(levels) == 1:
                ctx.sentinel_to_nan(grid, SENTINEL)
            outs.append(grid)
            self.last_stats.append({"levels": levels, "queries": queries})
        return torch.stack(outs, 0).to(latents.dtype)


def flash_topk_budget(M: int) -> int:
    """reference attention_processors.py:40-45."""
    if M == 3072:
        return 1024
    if M == 512:
        return 256
    return M // 3


class FlashVDMVolumeDecoding:
    """FlashVDM decoding with adaptive KV selection (reference volume_decoders.py:280-435) ->
    latents.dtype [B,N',N',N'] (N' = 381 for octree 384), NaN = unvisited.

    ``topk_mode`` 'mean' = FlashVDMCrossAttentionProcessor, 'merge' = FlashVDMTopMCrossAttentionProcessor
    (attention_processors.py:35-96).  No state is kept on ``geo_decoder`` (the reference installs a
    stateful processor on the shared module); refined queries are ordered by the STABLE
    sort of their bin id (the reference's ``index.sort()`` is unstable).  ``num_chunks``
    only changes how the reference batches whole bins / mini-grids, never the result: ignored."""

    def __init__(self, topk_mode='mean', keep_levels: bool = False):
        if topk_mode not in ['mean', 'merge']:
            raise ValueError(f'Unsupported topk_mode {topk_mode}, available: {["mean", "merge"]}')
        self.topk_mode = topk_mode
        self.keep_levels = keep_levels        # diagnostics: every level's grid of the last item (sentinel form) in last_levels
        self.last_levels = []

    @torch.no_grad()
    def __call__(self, latents, geo_decoder, bounds=1.01, num_chunks=10000, mc_level=0.0, octree_resolution=None,
                 min_resolution=63, mini_grid_num=4, enable_pbar=True, **kwargs):
        ctx = bind(latents, geo_decoder)
        dev = latents.device
        levels = flash_levels(octree_resolution, min_resolution, mini_grid_num)
        b6 = normalize_bounds(bounds)
        bbox_min, bbox_size = b6[:3], b6[3:] - b6[:3]
        merge = self.topk_mode == 'merge'
        outs = []
        self.last_stats = []
        for b in range(latents.shape[0]):
            ctx.prepare_kv(latents[b])
            T = flash_topk_budget(latents.shape[1])
            # ---- level 0: mini_grid_num^3 mini-grids, each with its own top-k tokens (reference :343-371)
            N0 = levels[0] + 1
            G0 = mini_grid_num ** 3
            pidx, tile_group, sidx, soff = ctx.flash_layout_minigrids(N0, mini_grid_num, 100)
            axes = axis_tables(bounds, levels[0])
            ctx.flash_select(sidx, (N0, N0, N0), soff, G0, T, False, axes=axes)      # level 0 is 'mean' in both modes
            grid = torch.empty((N0, N0, N0), dtype=torch.float32, device=dev)
            ctx.decode_flash(pidx, (N0, N0, N0), tile_group, grid, axes=axes)
            queries = [N0 ** 3]
            kept = [grid]
            # ---- refined levels: 6^3 spatial bins of the active queries (reference :373-431)
            for r in levels[1:]:
                n = r + 1
                index = refine_level(ctx, grid, mc_level, last=(r == levels[-1]), nf=n)
                cell = (bbox_size / r).astype(np.float32)
                bmin32 = bbox_min.astype(np.float32)
                nxt = torch.empty((n, n, n), dtype=torch.float32, device=dev)
                ctx.fill(nxt, NAN if r == levels[-1] and not self.keep_levels else SENTINEL)     # see HierarchicalVolumeDecoding
                nq = int(index.numel())
                queries.append(nq)
                if nq:
                    # bin ids (reference :394-403, float32 op for op), stable sort, per-bin padding and samples: one device pass
                    pidx, tile_group, sidx, soff = ctx.flash_layout_bins(index, (n, n, n), cell, bmin32, 30 if merge else 50)
                    ctx.flash_select(sidx, (n, n, n), soff, 216, T, merge, cell=cell, bmin=bmin32)
                    ctx.decode_flash(pidx, (n, n, n), tile_group, nxt, cell=cell, bmin=bmin32)
                grid = nxt
                kept.append(grid)
            if self.keep_levels:
                self.last_levels = kept[:-1] + [kept[-1].clone()]
            if self.keep_levels or len(levels) == 1:
                ctx.sentinel_to_nan(grid, SENTINEL)
            outs.append(grid)
            self.last_stats.append({"levels": levels, "queries": queries})
        return torch.stack(outs, 0).to(latents.dtype)
