"""ShapeVAE geometry-path configuration and synthetic weights.

The reference builds its weights with ``ShapeVAE(**params)`` (reference
``hy3dgen/shapegen/models/autoencoders/model.py:133-184``) and loads real
checkpoints through ``load_state_dict``.  This module knows the *names and
shapes* of that ``state_dict`` so that

* a live reference ``ShapeVAE`` / ``CrossAttentionDecoder`` can be consumed
  unchanged (``decoder_tensors_from_module``), and
* benchmarks and tests that run where the reference tree does not exist (the
  GPU box) can create random-init weights of the same architecture
  (``synthetic_state_dict``); the result loads into the reference class with
  ``strict=True`` (checked by ``oracle/make_golden.py``).

Nothing here computes; it only names tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Mapping

import torch


@dataclass(frozen=True)
class ShapeVAEConfig:
    """Constructor arguments of the reference ``ShapeVAE`` that shape the hot path
    (reference model.py:133-150).  Defaults are the Hunyuan3D-2 checkpoints'
    (reference project/image3d/shape.py:32-46)."""
    num_latents: int = 3072
    embed_dim: int = 64
    width: int = 1024
    heads: int = 16
    num_decoder_layers: int = 16
    geo_decoder_downsample_ratio: int = 1
    geo_decoder_mlp_expand_ratio: int = 4
    geo_decoder_ln_post: bool = True
    num_freqs: int = 8
    include_pi: bool = False
    qkv_bias: bool = False
    qk_norm: bool = True

    # ---- derived quantities for the geometry decoder (model.py:169-181) ----
    @property
    def dec_width(self) -> int:
        return self.width // self.geo_decoder_downsample_ratio

    @property
    def dec_heads(self) -> int:
        return self.heads // self.geo_decoder_downsample_ratio

    @property
    def dec_qk_norm(self) -> bool:
        # attention_blocks.py:460-461: no ln_post => no q/k norm
        return self.qk_norm and self.geo_decoder_ln_post

    @property
    def fourier_dim(self) -> int:
        return 3 * (2 * self.num_freqs + 1)

    def as_kwargs(self) -> dict:
        return asdict(self)


FULL = ShapeVAEConfig(num_latents=3072)
MINI = ShapeVAEConfig(num_latents=512)
# Turbo VAE variants: runtime knobs per SURVEY §8 (ratio 2 / expand 1 / no ln_post).
MINI_TURBO = ShapeVAEConfig(num_latents=512, geo_decoder_downsample_ratio=2,
                            geo_decoder_mlp_expand_ratio=1, geo_decoder_ln_post=False)


def fourier_frequencies(cfg: ShapeVAEConfig) -> torch.Tensor:
    """attention_blocks.py:84-98 (logspace=True branch)."""
    f = 2.0 ** torch.arange(cfg.num_freqs, dtype=torch.float32)
    if cfg.include_pi:
        f = f * torch.pi
    return f


def _linear(gen, out_f, in_f, bias=True, prefix=""):
    """PyTorch-default-like nn.Linear init: U(-1/sqrt(in), 1/sqrt(in))."""
    b = 1.0 / math.sqrt(in_f)
    d = {prefix + "weight": (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * b}
    if bias:
        d[prefix + "bias"] = (torch.rand(out_f, generator=gen) * 2 - 1) * b
    return d


def _ln(gen, n, prefix, jitter):
    """LayerNorm affine.  jitter>0 perturbs the default (1, 0) so that a wrong
    gamma/beta wiring cannot hide behind identity parameters."""
    w = torch.ones(n)
    b = torch.zeros(n)
    if jitter > 0:
        w = w + jitter * (torch.rand(n, generator=gen) * 2 - 1)
        b = b + jitter * (torch.rand(n, generator=gen) * 2 - 1)
    return {prefix + "weight": w, prefix + "bias": b}


def synthetic_state_dict(cfg: ShapeVAEConfig, seed: int = 0, ln_jitter: float = 0.1,
                         with_transformer: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init fp32 ``state_dict`` with the reference's key names and shapes.

    Keys follow SURVEY Appendix A.3 exactly, so ``ShapeVAE(**cfg).load_state_dict(sd)``
    succeeds strictly.
    """
    g = torch.Generator().manual_seed(seed)
    W, H = cfg.width, cfg.heads
    sd: Dict[str, torch.Tensor] = {}
    sd.update(_linear(g, W, cfg.embed_dim, prefix="post_kl."))
    if with_transformer:
        for i in range(cfg.num_decoder_layers):
            p = f"transformer.resblocks.{i}."
            sd.update(_linear(g, 3 * W, W, bias=cfg.qkv_bias, prefix=p + "attn.c_qkv."))
            sd.update(_linear(g, W, W, prefix=p + "attn.c_proj."))
            if cfg.qk_norm:
                sd.update(_ln(g, W // H, p + "attn.attention.q_norm.", ln_jitter))
                sd.update(_ln(g, W // H, p + "attn.attention.k_norm.", ln_jitter))
            sd.update(_ln(g, W, p + "ln_1.", ln_jitter))
            sd.update(_linear(g, 4 * W, W, prefix=p + "mlp.c_fc."))
            sd.update(_linear(g, W, 4 * W, prefix=p + "mlp.c_proj."))
            sd.update(_ln(g, W, p + "ln_2.", ln_jitter))
    Wd, Hd, r = cfg.dec_width, cfg.dec_heads, cfg.geo_decoder_mlp_expand_ratio
    p = "geo_decoder."
    sd.update(_linear(g, Wd, cfg.fourier_dim, prefix=p + "query_proj."))
    if cfg.geo_decoder_downsample_ratio != 1:
        sd.update(_linear(g, Wd, W, prefix=p + "latents_proj."))
    c = p + "cross_attn_decoder."
    sd.update(_linear(g, Wd, Wd, bias=cfg.qkv_bias, prefix=c + "attn.c_q."))
    sd.update(_linear(g, 2 * Wd, Wd, bias=cfg.qkv_bias, prefix=c + "attn.c_kv."))
    sd.update(_linear(g, Wd, Wd, prefix=c + "attn.c_proj."))
    if cfg.dec_qk_norm:
        sd.update(_ln(g, Wd // Hd, c + "attn.attention.q_norm.", ln_jitter))
        sd.update(_ln(g, Wd // Hd, c + "attn.attention.k_norm.", ln_jitter))
    sd.update(_ln(g, Wd, c + "ln_1.", ln_jitter))
    sd.update(_ln(g, Wd, c + "ln_2.", ln_jitter))
    sd.update(_ln(g, Wd, c + "ln_3.", ln_jitter))
    sd.update(_linear(g, r * Wd, Wd, prefix=c + "mlp.c_fc."))
    sd.update(_linear(g, Wd, r * Wd, prefix=c + "mlp.c_proj."))
    if cfg.geo_decoder_ln_post:
        sd.update(_ln(g, Wd, p + "ln_post.", ln_jitter))
    sd.update(_linear(g, 1, Wd, prefix=p + "output_proj."))
    return sd


def sparsify_field(sd: Dict[str, torch.Tensor], cfg: ShapeVAEConfig, keep_freqs: int = 2,
                   gain: float = 1.0, bias: float = 0.0) -> Dict[str, torch.Tensor]:
    """Deterministic post-edit of SURVEY §8(d): drop the high Fourier frequencies
    from ``query_proj`` and rescale/shift the output head so that the random
    field is low-frequency and saturating (sparse near-surface set)."""
    sd = dict(sd)
    w = sd["geo_decoder.query_proj.weight"].clone()
    F = cfg.num_freqs
    for a in range(3):
        for f in range(keep_freqs, F):
            w[:, 3 + F * a + f] = 0.0
            w[:, 3 + 3 * F + F * a + f] = 0.0
    sd["geo_decoder.query_proj.weight"] = w
    sd["geo_decoder.output_proj.weight"] = sd["geo_decoder.output_proj.weight"] * gain
    sd["geo_decoder.output_proj.bias"] = sd["geo_decoder.output_proj.bias"] * gain + bias
    return sd


def synthetic_latents(cfg: ShapeVAEConfig, batch: int = 1, seed: int = 1234) -> torch.Tensor:
    """``z ~ N(0,1)`` of shape ``[B, num_latents, embed_dim]`` (what
    ``prepare_latents`` yields, reference pipelines.py:470-485); item b uses
    seed ``seed + b``."""
    zs = [torch.randn(cfg.num_latents, cfg.embed_dim,
                      generator=torch.Generator().manual_seed(seed + b)) for b in range(batch)]
    return torch.stack(zs, 0)


GEO_PREFIX = "geo_decoder."


def geo_decoder_state(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Sub-dict of the ``geo_decoder.*`` entries with the prefix stripped — the
    same keys ``CrossAttentionDecoder.state_dict()`` yields."""
    return {k[len(GEO_PREFIX):]: v for k, v in sd.items() if k.startswith(GEO_PREFIX)}


def config_from_geo_decoder(geo_decoder) -> ShapeVAEConfig:
    """Read the hyper-parameters the kernels need from a live reference
    ``CrossAttentionDecoder`` (attention_blocks.py:435-476) by attribute, never
    by calling it."""
    sd = geo_decoder.state_dict()
    fe = geo_decoder.fourier_embedder
    Wd = sd["query_proj.weight"].shape[0]
    ratio = int(getattr(geo_decoder, "downsample_ratio", 1))
    heads = int(geo_decoder.cross_attn_decoder.attn.heads)
    freqs = fe.frequencies.detach().float().cpu()
    include_pi = bool(abs(float(freqs[0]) - math.pi) < 1e-4)
    if not bool(fe.include_input):
        raise TypeError("FourierEmbedder(include_input=False) is not a ShapeVAE configuration")
    return ShapeVAEConfig(
        num_latents=0, embed_dim=0, width=Wd * ratio, heads=heads * ratio, num_decoder_layers=0,
        geo_decoder_downsample_ratio=ratio,
        geo_decoder_mlp_expand_ratio=sd["cross_attn_decoder.mlp.c_fc.weight"].shape[0] // Wd,
        geo_decoder_ln_post=bool(geo_decoder.enable_ln_post),
        num_freqs=int(fe.num_freqs), include_pi=include_pi,
        qkv_bias="cross_attn_decoder.attn.c_q.bias" in sd,
        qk_norm="cross_attn_decoder.attn.attention.q_norm.weight" in sd,
    )
