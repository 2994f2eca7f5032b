"""CPU, build container only (skipped where /root/reference is absent): the drop-in boundary against
the LIVE reference classes — strict state_dict compatibility and slot swapping by hy3dgeo.install."""
import pytest
import torch

import hy3dgeo
from hy3dgeo import weights as W
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ns():
    return ref_loader.load()


def test_state_dict_loads_strictly_and_config_is_recovered(ns):
    for cfg in (W.MINI_TURBO,):
        sd = W.synthetic_state_dict(cfg, seed=0)
        vae = ref_loader.build_shapevae(ns, cfg, sd)               # strict=True inside
        back = W.config_from_geo_decoder(vae.geo_decoder)
        assert (back.dec_width, back.dec_heads, back.geo_decoder_mlp_expand_ratio, back.geo_decoder_ln_post,
                back.dec_qk_norm, back.num_freqs, back.include_pi) == \
               (cfg.dec_width, cfg.dec_heads, cfg.geo_decoder_mlp_expand_ratio, cfg.geo_decoder_ln_post,
                cfg.dec_qk_norm, cfg.num_freqs, cfg.include_pi)


def test_install_swaps_plugin_slots(ns):
    cfg = W.MINI_TURBO
    vae = ref_loader.build_shapevae(ns, cfg, W.synthetic_state_dict(cfg, seed=0))
    assert type(vae.volume_decoder).__module__.startswith("hy3dgen")
    hy3dgeo.install(vae)
    assert isinstance(vae.volume_decoder, hy3dgeo.VanillaVolumeDecoder)
    assert isinstance(vae.surface_extractor, hy3dgeo.MCSurfaceExtractor)
    vae.enable_flashvdm_decoder(enabled=True, adaptive_kv_selection=True, topk_mode="merge", mc_algo="mc")
    assert isinstance(vae.volume_decoder, hy3dgeo.FlashVDMVolumeDecoding) and vae.volume_decoder.topk_mode == "merge"
    vae.enable_flashvdm_decoder(enabled=True, adaptive_kv_selection=False, mc_algo="mc")
    assert isinstance(vae.volume_decoder, hy3dgeo.HierarchicalVolumeDecoding)
    with pytest.raises(ValueError):
        vae.enable_flashvdm_decoder(mc_algo="bogus")
    # reference FlashVDM instance -> hy3dgeo instance with the same mode
    vae.volume_decoder = ns.vd.FlashVDMVolumeDecoding("merge")
    hy3dgeo.install(vae)
    assert isinstance(vae.volume_decoder, hy3dgeo.FlashVDMVolumeDecoding) and vae.volume_decoder.topk_mode == "merge"
    # no CPU path: latents2mesh on CPU latents must fail loudly, not fall back to the reference
    with pytest.raises(RuntimeError):
        vae.latents2mesh(torch.zeros(1, 512, 1024), bounds=1.01, mc_level=0.0, num_chunks=8000, octree_resolution=16,
                         mc_algo="mc", enable_pbar=False)
