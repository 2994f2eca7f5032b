"""ORACLE (test infrastructure, never shipped): CPU fp32 restatement of the
reference's geometry decoder arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path under
``hunyuan3d-2_b200/`` never does.

Pinned against the real reference classes by ``oracle/make_golden.py`` (run in
the build container where ``/root/reference`` exists); the resulting vectors
live in ``tests/golden/`` and ``tests/test_oracle_golden.py`` re-checks this
file against them everywhere (incl. the GPU box, where the reference is absent).

Every function cites the reference lines it follows (paths relative to
``/root/reference/hy3dgen/shapegen/models/autoencoders/``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def fourier_embed(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """attention_blocks.py:112-130 with include_input=True:
    ``[x, sin(x (x) f), cos(x (x) f)]``; the outer product is flattened
    axis-major (x*f0..x*f7, y*f0.., z*f0..)."""
    emb = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat((x, emb.sin(), emb.cos()), dim=-1)


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def project_latents(sd: Dict[str, torch.Tensor], latents: torch.Tensor) -> torch.Tensor:
    """attention_blocks.py:487-488 (only when downsample_ratio != 1)."""
    if "latents_proj.weight" in sd:
        latents = _lin(latents, sd, "latents_proj")
    return latents


def kv_heads(sd, latents, heads):
    """K/V per head from the latent tokens: ln_2 (attention_blocks.py:296), c_kv
    (:257), per-head split ``[k|v]`` and k_norm (:205-211).  Returns
    ``k, v`` of shape ``[B, H, M, d]``."""
    c = "cross_attn_decoder."
    data = _ln(project_latents(sd, latents), sd, c + "ln_2", 1e-6)
    kv = _lin(data, sd, c + "attn.c_kv")
    B, M, W2 = kv.shape
    d = W2 // heads // 2
    kv = kv.view(B, M, heads, 2 * d)
    k, v = kv[..., :d], kv[..., d:]
    if c + "attn.attention.k_norm.weight" in sd:
        k = _ln(k, sd, c + "attn.attention.k_norm", 1e-6)
    return k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)


def q_heads(sd, x0, heads):
    """ln_1 (attention_blocks.py:296), c_q (:250), per-head view and q_norm
    (:204,210).  ``x0`` is ``[B, P, W]``; returns ``[B, H, P, d]``."""
    c = "cross_attn_decoder."
    q = _lin(_ln(x0, sd, c + "ln_1", 1e-6), sd, c + "attn.c_q")
    B, P, W = q.shape
    q = q.view(B, P, heads, W // heads)
    if c + "attn.attention.q_norm.weight" in sd:
        q = _ln(q, sd, c + "attn.attention.q_norm", 1e-6)
    return q.permute(0, 2, 1, 3)


def sdpa(q, k, v):
    """torch SDPA semantics (attention_processors.py:29-32): softmax(q k^T / sqrt(d)) v."""
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    return torch.softmax(s, dim=-1) @ v


def decoder_tail(sd, x0, attn, *, ln_post: bool):
    """c_proj + residual, ln_3/MLP(GELU erf) + residual, [ln_post], output_proj
    (attention_blocks.py:259, 296-297, 175-181, 490-492).  ``attn`` is
    ``[B, H, P, d]``."""
    c = "cross_attn_decoder."
    B, H, P, d = attn.shape
    a = attn.permute(0, 2, 1, 3).reshape(B, P, H * d)
    x1 = x0 + _lin(a, sd, c + "attn.c_proj")
    h = F.gelu(_lin(_ln(x1, sd, c + "ln_3", 1e-6), sd, c + "mlp.c_fc"))
    x2 = x1 + _lin(h, sd, c + "mlp.c_proj")
    if ln_post:
        x2 = _ln(x2, sd, "ln_post", 1e-5)
    return _lin(x2, sd, "output_proj")


def geo_decoder_forward(sd: Dict[str, torch.Tensor], queries: torch.Tensor, latents: torch.Tensor,
                        freqs: torch.Tensor, heads: int, kv_select=None) -> torch.Tensor:
    """``CrossAttentionDecoder.forward`` (attention_blocks.py:483-493).

    sd       : fp32 state dict of the decoder (keys without ``geo_decoder.``)
    queries  : [B, P, 3];  latents : [B, M, W]
    kv_select: optional callable ``(q, k, v) -> attention output`` standing in
               for a FlashVDM processor (attention_processors.py:35-96); default
               is plain SDPA.
    returns  : logits [B, P, 1]
    """
    ln_post = "ln_post.weight" in sd
    x0 = _lin(fourier_embed(queries, freqs), sd, "query_proj")
    k, v = kv_heads(sd, latents, heads)
    q = q_heads(sd, x0, heads)
    attn = sdpa(q, k, v) if kv_select is None else kv_select(q, k, v)
    return decoder_tail(sd, x0, attn, ln_post=ln_post)


# ----------------------------------------------------------------------------
# FlashVDM KV selection (attention_processors.py:35-96)
# ----------------------------------------------------------------------------

def flash_topk_budget(M: int) -> int:
    """attention_processors.py:40-45."""
    if M == 3072:
        return 1024
    if M == 512:
        return 256
    return M // 3


def mean_similarity(q_chunk, k, stride):
    """``mean over sampled queries of (q . k)``, unscaled (attention_processors.py:48-50,
    74-76).  q_chunk [B,H,P,d] -> sim [B,H,M]."""
    q1 = q_chunk[:, :, ::stride, :]
    return (q1 @ k.transpose(-1, -2)).mean(-2)


def select_mean(q_chunk, k, v, topk, stride):
    """'mean' mode: per (batch, head) keep the ``topk`` tokens of largest mean
    similarity (attention_processors.py:47-55, 73-81).  Returns gathered k, v and
    the selected ids [B,H,topk]."""
    sim = mean_similarity(q_chunk, k, stride)
    ids = torch.topk(sim, dim=-1, k=topk).indices
    g = ids[..., None].expand(-1, -1, -1, v.shape[-1])
    return torch.gather(k, -2, g), torch.gather(v, -2, g), ids


def select_merge(q_chunk, k, v, stride=30, thresh=1e-6):
    """'merge' mode (attention_processors.py:84-96): stride-30 queries, unscaled
    softmax over tokens, mean over heads, tokens with any probability > 1e-6;
    one token set shared by all heads.  Batch must be 1 (as in the reference,
    which indexes ``where(...)[2]``)."""
    q1 = q_chunk[:, :, ::stride, :]
    sim = (q1 @ k.transpose(-1, -2)).softmax(-1).mean(1)         # [B, P', M]
    ids = torch.unique(torch.where(sim > thresh)[2])
    g = ids.view(1, 1, -1, 1).expand(-1, v.shape[1], -1, v.shape[-1])
    return torch.gather(k, -2, g), torch.gather(v, -2, g), ids


class FlashProcessorOracle:
    """Stateful stand-in for ``FlashVDMCrossAttentionProcessor`` /
    ``FlashVDMTopMCrossAttentionProcessor``: set ``.topk`` to True (level 0),
    False (plain) or ``(ids, counts)`` (refined levels) before each decoder
    call, exactly like volume_decoders.py:362,415,424 do."""

    def __init__(self, mode="mean"):
        assert mode in ("mean", "merge")
        self.mode = mode
        self.topk = False
        self.last_selection = []      # list of id tensors, for selection-parity tests

    def __call__(self, q, k, v):
        T = flash_topk_budget(k.shape[-2])
        self.last_selection = []
        if self.topk is True:
            k0, v0, ids = select_mean(q, k, v, T, 100)          # level 0 is 'mean' in both modes (:47-55)
            self.last_selection.append(ids)
            out = sdpa(q, k0, v0)
        elif self.topk is False:
            out = sdpa(q, k, v)
        else:
            _, counts = self.topk
            outs, start = [], 0
            for cnt in counts:
                qc = q[:, :, start:start + cnt, :]
                if self.mode == "mean":
                    k0, v0, ids = select_mean(qc, k, v, T, 50)
                else:
                    k0, v0, ids = select_merge(qc, k, v)
                self.last_selection.append(ids)
                outs.append(sdpa(qc, k0, v0))
                start += cnt
            out = torch.cat(outs, dim=-2)
        self.topk = False
        return out


# ----------------------------------------------------------------------------
# ShapeVAE.forward (model.py:186-189) : post_kl + 16 pre-LN self-attention blocks
# ----------------------------------------------------------------------------

def shapevae_forward(sd: Dict[str, torch.Tensor], z: torch.Tensor, heads: int) -> torch.Tensor:
    """post_kl (model.py:187) then ``Transformer`` (attention_blocks.py:429-432);
    block = ``x += c_proj(SDPA(q_norm q, k_norm k, v)); x += MLP(ln_2 x)`` (:391-394);
    c_qkv output viewed [B,n,H,3d] and split [q|k|v] per head (:318-321)."""
    x = F.linear(z, sd["post_kl.weight"], sd["post_kl.bias"])
    i = 0
    while f"transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"transformer.resblocks.{i}."
        y = _ln(x, sd, p + "ln_1", 1e-6)
        qkv = _lin(y, sd, p + "attn.c_qkv")
        B, n, W3 = qkv.shape
        d = W3 // heads // 3
        qkv = qkv.view(B, n, heads, 3 * d)
        q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        if p + "attn.attention.q_norm.weight" in sd:
            q = _ln(q, sd, p + "attn.attention.q_norm", 1e-6)
            k = _ln(k, sd, p + "attn.attention.k_norm", 1e-6)
        o = sdpa(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3))
        o = o.permute(0, 2, 1, 3).reshape(B, n, heads * d)
        x = x + _lin(o, sd, p + "attn.c_proj")
        h = F.gelu(_lin(_ln(x, sd, p + "ln_2", 1e-6), sd, p + "mlp.c_fc"))
        x = x + _lin(h, sd, p + "mlp.c_proj")
        i += 1
    return x
