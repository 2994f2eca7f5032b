"""ShapeVAE geometry path — host-side mirror of the reference ``VectsetVAE`` /
``ShapeVAE`` plugin wiring (``hy3dgen/shapegen/models/autoencoders/model.py``).

Two ways in:

* ``install(vae)`` swaps the two plugin slots of a live reference ``ShapeVAE``
  (``vae.volume_decoder`` / ``vae.surface_extractor``, reference model.py:102-103)
  for the B200 implementations; ``Hunyuan3DDiTFlowMatchingPipeline`` then runs
  unchanged.
* ``B200ShapeVAE`` is a self-contained holder with the same ``state_dict`` keys,
  ``forward`` and ``latents2mesh`` for places where the reference package is not
  importable (benchmarks, the GPU box).
"""
from __future__ import annotations

import types
from typing import Dict

import torch
import torch.nn.functional as F

from . import weights as W
from .surface_extractors import MCSurfaceExtractor, SurfaceExtractors
from .utils import synchronize_timer
from .volume_decoders import FlashVDMVolumeDecoding, HierarchicalVolumeDecoding, VanillaVolumeDecoder


class GeoDecoder:
    """Weights of ``CrossAttentionDecoder`` with the attribute surface the volume
    decoders read (reference attention_blocks.py:435-476).  Not callable: queries
    are evaluated by the CUDA kernels only."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: W.ShapeVAEConfig):
        self._sd = dict(sd)
        self.enable_ln_post = cfg.geo_decoder_ln_post
        self.downsample_ratio = cfg.geo_decoder_downsample_ratio
        self.fourier_embedder = types.SimpleNamespace(
            frequencies=W.fourier_frequencies(cfg), num_freqs=cfg.num_freqs, include_input=True)
        self.cross_attn_decoder = types.SimpleNamespace(attn=types.SimpleNamespace(heads=cfg.dec_heads))
        self.count = 0

    def state_dict(self):
        return self._sd

    def set_cross_attention_processor(self, processor):     # reference :477-478; selection is a decoder mode here
        pass


class B200ShapeVAE:
    """Mirror of ``ShapeVAE`` (reference model.py:132-189) restricted to decoding."""

    def __init__(self, cfg: W.ShapeVAEConfig, state_dict: Dict[str, torch.Tensor], device="cuda",
                 volume_decoder=None, surface_extractor=None, scale_factor: float = 1.0):
        self.cfg = cfg
        self.device = torch.device(device)
        self.sd = {k: v.to(self.device, torch.float32) for k, v in state_dict.items()}
        self.geo_decoder = GeoDecoder(W.geo_decoder_state(self.sd), cfg)
        self.volume_decoder = volume_decoder if volume_decoder is not None else VanillaVolumeDecoder()
        self.surface_extractor = surface_extractor if surface_extractor is not None else MCSurfaceExtractor()
        self.scale_factor = scale_factor
        self.latent_shape = (cfg.num_latents, cfg.embed_dim)

    def _tc_ok(self, latents) -> bool:
        c = self.cfg
        return (self.device.type == "cuda" and latents.shape[-2] % 128 == 0 and c.width % 256 == 0 and c.width // c.heads == 64
                and c.embed_dim % 64 == 0 and c.num_decoder_layers > 0)

    @torch.no_grad()
    def forward(self, latents: torch.Tensor, dtype=torch.float32, impl: str = None, group=None) -> torch.Tensor:
        """ShapeVAE.forward (reference model.py:186-189): post_kl + Transformer.

        impl='tc' (the default and the product path): hand-written tcgen05 kernels of libhy3dgeo.so
        (``hy3d_transformer_forward``: 3-term split fp16 GEMMs = fp32-grade, LayerNorms folded, fp16
        self-attention), float32 result.  Shapes those kernels do not tile (head dim != 64, width not a multiple
        of 256, token count not a multiple of 128) raise — there is no silent library fallback.
        ``group`` (a torch.distributed process group, or True for the default group): the latents are known on every rank
        and the pass runs sequence-parallel over the group (hy3dgeo.h: hy3d_transformer_begin / layer_kv / layer_rest / end);
        every rank gets the full result.
        impl='torch' must be asked for explicitly: the same network on cuBLAS + SDPA library ops in ``dtype``, kept
        as a cross-check for the parity tests (47 ms in fp32 for 3072 tokens vs ~5 ms)."""
        if impl is None:
            impl = "tc"
        if impl not in ("tc", "torch"):
            raise ValueError(f"impl must be 'tc' or 'torch', got {impl!r}")
        if impl == "tc":
            if not self._tc_ok(latents):
                raise RuntimeError(
                    "B200ShapeVAE.forward: the tcgen05 latent transformer needs head_dim 64, width % 256 == 0, embed_dim % 64 == 0 "
                    f"and a token count that is a multiple of 128 (got width {self.cfg.width}, heads {self.cfg.heads}, embed_dim "
                    f"{self.cfg.embed_dim}, tokens {latents.shape[-2]}); pass impl='torch' explicitly for the library statement")
            from ._lib import get_context
            ctx = get_context(self.device)
            ctx.set_transformer(self.sd, self.cfg, key=id(self.sd), owner=self)
            z = latents.to(self.device)
            if group is not None:
                # sequence parallel over a process group: each rank runs M / world token rows, the K / V tile images are
                # all-gathered once per layer (what a rank repeats for every latent does not shrink with more GPUs otherwise)
                import torch.distributed as dist
                world = dist.get_world_size(group if group is not True else None)
                if world > 1 and z.shape[-2] % (128 * world) == 0:
                    g = None if group is True else group
                    return torch.stack([ctx.transformer_forward_parallel(z[b], self.cfg.heads, self.cfg.num_decoder_layers, g)
                                        for b in range(z.shape[0])], 0)
            return torch.stack([ctx.transformer_forward(z[b]) for b in range(z.shape[0])], 0)
        return self._forward_torch(latents, dtype)

    # Library GEMM/SDPA statement of the same network (cuBLAS + torch SDPA).
    @torch.no_grad()
    def _forward_torch(self, latents: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        sd, H = self.sd, self.cfg.heads
        x = F.linear(latents.to(self.device, dtype), sd["post_kl.weight"].to(dtype), sd["post_kl.bias"].to(dtype))
        for i in range(self.cfg.num_decoder_layers):
            p = f"transformer.resblocks.{i}."
            g = lambda n: sd[p + n].to(dtype) if (p + n) in sd else None
            y = F.layer_norm(x, (x.shape[-1],), g("ln_1.weight"), g("ln_1.bias"), 1e-6)
            qkv = F.linear(y, g("attn.c_qkv.weight"), g("attn.c_qkv.bias"))
            B, n, W3 = qkv.shape
            d = W3 // H // 3
            qkv = qkv.view(B, n, H, 3 * d)
            q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
            if g("attn.attention.q_norm.weight") is not None:
                q = F.layer_norm(q, (d,), g("attn.attention.q_norm.weight"), g("attn.attention.q_norm.bias"), 1e-6)
                k = F.layer_norm(k, (d,), g("attn.attention.k_norm.weight"), g("attn.attention.k_norm.bias"), 1e-6)
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
            o = o.transpose(1, 2).reshape(B, n, H * d)
            x = x + F.linear(o, g("attn.c_proj.weight"), g("attn.c_proj.bias"))
            h = F.gelu(F.linear(F.layer_norm(x, (x.shape[-1],), g("ln_2.weight"), g("ln_2.bias"), 1e-6),
                                g("mlp.c_fc.weight"), g("mlp.c_fc.bias")))
            x = x + F.linear(h, g("mlp.c_proj.weight"), g("mlp.c_proj.bias"))
        return x

    __call__ = forward

    def latents2mesh(self, latents: torch.Tensor, **kwargs):
        """reference model.py:105-110 (same timer names so HY3DGEN_DEBUG logs line up)."""
        with synchronize_timer('Volume decoding'):
            grid_logits = self.volume_decoder(latents, self.geo_decoder, **kwargs)
        with synchronize_timer('Surface extraction'):
            outputs = self.surface_extractor(grid_logits, **kwargs)
        return outputs

    def enable_flashvdm_decoder(self, enabled: bool = True, adaptive_kv_selection=True, topk_mode='mean', mc_algo='dmc'):
        enable_flashvdm_decoder(self, enabled, adaptive_kv_selection, topk_mode, mc_algo)


def enable_flashvdm_decoder(vae, enabled: bool = True, adaptive_kv_selection=True, topk_mode='mean', mc_algo='dmc'):
    """reference model.py:112-129 with the B200 classes."""
    if enabled:
        if adaptive_kv_selection:
            vae.volume_decoder = FlashVDMVolumeDecoding(topk_mode)
        else:
            vae.volume_decoder = HierarchicalVolumeDecoding()
        if mc_algo not in SurfaceExtractors.keys():
            raise ValueError(f'Unsupported mc_algo {mc_algo}, available: {list(SurfaceExtractors.keys())}')
        vae.surface_extractor = SurfaceExtractors[mc_algo]()
    else:
        vae.volume_decoder = VanillaVolumeDecoder()
        vae.surface_extractor = MCSurfaceExtractor()


_REF_TO_B200 = {
    "VanillaVolumeDecoder": VanillaVolumeDecoder,
    "HierarchicalVolumeDecoding": HierarchicalVolumeDecoding,
    "FlashVDMVolumeDecoding": FlashVDMVolumeDecoding,
}


def install(vae):
    """Swap the plugin slots of a live reference ``VectsetVAE`` (model.py:102-103) for their
    B200 counterparts, keeping whichever decoder kind is currently selected, and rebind
    ``enable_flashvdm_decoder`` so later switches stay on the B200 classes.  Returns ``vae``."""
    kind = type(vae.volume_decoder).__name__.replace("Patched", "")
    if kind not in _REF_TO_B200:
        raise TypeError(f"unknown volume decoder {kind}")
    if kind == "FlashVDMVolumeDecoding":
        mode = 'merge' if 'TopM' in type(getattr(vae.volume_decoder, 'processor', None)).__name__ else 'mean'
        vae.volume_decoder = FlashVDMVolumeDecoding(mode)
    else:
        vae.volume_decoder = _REF_TO_B200[kind]()
    ext = type(vae.surface_extractor).__name__
    if ext == "MCSurfaceExtractor":
        vae.surface_extractor = MCSurfaceExtractor()
    vae.enable_flashvdm_decoder = types.MethodType(
        lambda self, enabled=True, adaptive_kv_selection=True, topk_mode='mean', mc_algo='dmc':
        enable_flashvdm_decoder(self, enabled, adaptive_kv_selection, topk_mode, mc_algo), vae)
    return vae
