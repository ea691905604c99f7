"""GPU: hub encoder families of SURVEY.md section 8(f) rank 4 on the DINOv2 / ViT kernels through the C ABI -- midnight ([class || mean
of patch tokens] head, Pillow BILINEAR preset, atlas_patch/models/patch/midnight.py), phikon_v1 (transformers ViTModel + uint8
bilinear-antialias processor, phikon.py:41-56), phikon_v2 (Dinov2Model ViT-L/16 + bicubic processor, phikon.py:90-105), hibou /
openmidnight (4 register tokens) and the CLIP image towers of plip / quilt (pre-LayerNorm, QuickGELU, visual projection):
preprocess pixels bit-exact vs the integer restatements (pinned on the CPU against the reference's own preprocess objects),
features within 1e-3 relative of transformers' models run in fp32 on the CPU with the same seeded weights."""
import numpy as np
import pytest

from oracle import hub_families as hf

pytestmark = pytest.mark.gpu


def _slide():
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    return SyntheticWSI(make_spec(4096, 3072, 11, mpp=0.5))


@pytest.mark.parametrize("name,P", [("midnight_test_tiny", 224), ("midnight_test_tiny", 256), ("midnight_test_tiny", 512),
                                    ("phikon_v1_test_tiny", 224), ("phikon_v1_test_tiny", 256), ("phikon_v1_test_tiny", 300),
                                    ("phikon_v2_test_tiny", 224), ("phikon_v2_test_tiny", 256),
                                    ("hibou_test_tiny", 224), ("hibou_test_tiny", 256),                  # 4 register tokens: 261-token sequence
                                    ("openmidnight_test_tiny", 224), ("openmidnight_test_tiny", 512), ("h_optimus_test_tiny", 256),
                                    ("pathorchestra_test_tiny", 256),                                    # timm keys, Pillow BILINEAR to 224
                                    ("prov_gigapath_test_tiny", 224), ("prov_gigapath_test_tiny", 256), ("prov_gigapath_test_tiny", 512),  # Pillow BICUBIC
                                    # CLIP towers: pre-LayerNorm, QuickGELU, 128-wide visual projection (ViT-B/32: 50 tokens; B/16: 197)
                                    ("plip_test_tiny", 224), ("plip_test_tiny", 256), ("quilt_b_16_test_tiny", 256),
                                    # OpenAI CLIP through open_clip: Pillow BICUBIC preprocess, open_clip key layout; L/14 = 257 tokens
                                    ("clip_vit_b_32_test_tiny", 224), ("clip_vit_l_14_test_tiny", 256)])
def test_tiny_family_pixels_bit_exact_and_features(name, P):
    import torch

    from atlaspatch_b200.encoder import FAMILY_RECIPES, B200FeatureExtractor
    from atlaspatch_b200.synthetic import render_region_host

    wsi = _slide()
    rng = np.random.default_rng(P + len(name))
    n = 9
    xy = np.stack([rng.integers(0, wsi.w - P, n), rng.integers(0, wsi.h - P, n)], 1)
    xy[-1] = (wsi.w - P // 2, wsi.h - P // 3)                                   # overhang: zero padding like IWSI.extract
    rows = np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
    patches = [render_region_host(wsi.spec, int(x), int(y), P, P) for x, y in xy]
    sd = hf.state_dict(name, seed=6)
    sd_in = sd
    if name.startswith("openmidnight"):     # the reference holds these weights in facebookresearch's key layout (openmidnight.py:49-63)
        from tests.test_oracle_hub_families import hf_to_fb_names

        sd_in = hf_to_fb_names(sd, 2, True)
    if name.startswith(("pathorchestra", "prov_gigapath")):     # timm's key layout with a class position (no_embed_class = False)
        from tests.test_oracle_hub_families import hf_to_fb_names

        sd_in = hf_to_fb_names(sd, 2, name.startswith("prov_gigapath"))
        sd_in.pop("mask_token")
        if name.startswith("prov_gigapath"):
            for i in range(2):
                for k in ("weight", "bias"):
                    sd_in[f"blocks.{i}.mlp.fc1.{k}"], sd_in[f"blocks.{i}.mlp.fc2.{k}"] = sd_in.pop(f"blocks.{i}.mlp.w12.{k}"), sd_in.pop(f"blocks.{i}.mlp.w3.{k}")
    if name.startswith("clip_vit"):         # open_clip's key layout (clip.py:36-40)
        from tests.test_oracle_hub_families import hf_to_openclip_names

        sd_in = hf_to_openclip_names(sd, 2)
    if name.startswith("h_optimus"):        # timm's key layout (hoptimus.py:53-58), no position on the class token
        from tests.test_oracle_hub_families import hf_to_timm_names

        sd["embeddings.position_embeddings"][:, 0] = 0.0
        sd_in = hf_to_timm_names(sd, 2, True)
    ext = B200FeatureExtractor(name, sd_in, input_patch=P, max_batch=4)         # 9 patches -> three forward chunks
    pool = FAMILY_RECIPES[name]["pool"]
    rows_dev = torch.from_numpy(rows).cuda()
    pix = ext.preprocess_pixels(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows_dev[-4:])
    for i in range(4):
        assert np.array_equal(pix[i], hf.pixels(name, patches[n - 4 + i])), i
    want = hf.extract_features(patches, sd, name)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows_dev).cpu().numpy()
    assert got.shape == want.shape == (n, ext.embedding_dim)
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    print(name, P, "rel err per row:", rel)
    assert rel.max() < 1e-3, rel
    if pool == 1:                                                               # both halves separately: the mean half has a much smaller norm
        d = got.shape[1] // 2
        for half in (slice(0, d), slice(d, 2 * d)):
            r = np.linalg.norm(got[:, half] - want[:, half], axis=1) / np.linalg.norm(want[:, half], axis=1)
            assert r.max() < 1e-3, r
    got_host = ext.extract_batch(patches, batch_size=4)                         # FeatureExtractor contract, host patches
    assert got_host.shape == got.shape and np.abs(got_host - got).max() < 1e-5
    assert ext.extract_batch([]).shape == (0, ext.embedding_dim)
    ext.cleanup()


def test_midnight_head_at_full_width():
    """ViT-g width (hidden 1536, 24 heads, SwiGLU 4096) with 2 layers: the 3072-float [class || mean] rows of midnight.py:53-61."""
    import torch

    from atlaspatch_b200 import dinov2 as d2
    from atlaspatch_b200 import weights as wt
    from atlaspatch_b200.encoder import FAMILY_RECIPES, B200FeatureExtractor
    from atlaspatch_b200.synthetic import render_region_host

    name = "midnight_test_wide"
    wt.DINOV2_SPECS[name] = (2, 24, 1536, True)
    d2.DINOV2_CONFIGS[name] = (14, 2, 24, 1536, 4096, True)
    FAMILY_RECIPES[name] = FAMILY_RECIPES["midnight"]
    try:
        wsi = _slide()
        P, n = 224, 5
        rng = np.random.default_rng(1)
        xy = np.stack([rng.integers(0, wsi.w - P, n), rng.integers(0, wsi.h - P, n)], 1)
        rows = np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
        patches = [render_region_host(wsi.spec, int(x), int(y), P, P) for x, y in xy]
        sd = hf.state_dict(name, seed=2)
        want = hf.extract_features(patches, sd, name)
        ext = B200FeatureExtractor(name, sd, input_patch=P, max_batch=4)
        got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, torch.from_numpy(rows).cuda()).cpu().numpy()
        assert got.shape == want.shape == (n, 3072)
        rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
        assert rel.max() < 1e-3, rel
        ext.cleanup()
    finally:
        wt.DINOV2_SPECS.pop(name, None)
        d2.DINOV2_CONFIGS.pop(name, None)
        FAMILY_RECIPES.pop(name, None)


def test_class_mean_head_skips_register_tokens():
    """[class || mean of the patch tokens] over a register-token sequence: virchow.py:110-114 `output[:, 5:]`, hoptimus.py:157-161
    `output[:, m.num_prefix_tokens:]` -- the mean must leave the 4 register rows out."""
    import torch
    from PIL import Image

    from atlaspatch_b200 import dinov2 as d2
    from atlaspatch_b200 import weights as wt
    from atlaspatch_b200.encoder import FAMILY_RECIPES, B200FeatureExtractor
    from atlaspatch_b200.synthetic import render_region_host

    name = "hibou_regpool_test"
    wt.DINOV2_SPECS[name], wt.DINOV2_REGISTERS[name] = (2, 6, 384, True), 4
    d2.DINOV2_CONFIGS[name], d2.DINOV2_REGISTERS[name] = (14, 2, 6, 384, 1024, True), 4
    FAMILY_RECIPES[name] = dict(FAMILY_RECIPES["hibou_test_tiny"], pool=1)
    try:
        wsi = _slide()
        P, n = 224, 6
        rng = np.random.default_rng(4)
        xy = np.stack([rng.integers(0, wsi.w - P, n), rng.integers(0, wsi.h - P, n)], 1)
        rows = np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
        patches = [render_region_host(wsi.spec, int(x), int(y), P, P) for x, y in xy]
        sd = hf.state_dict(name, seed=3)
        model, pre = hf.build_model(name, sd), hf.make_preprocess(name)
        with torch.inference_mode():
            h = model(pixel_values=torch.stack([pre(Image.fromarray(p)) for p in patches])).last_hidden_state
            want = torch.cat([h[:, 0], h[:, 5:].mean(1)], dim=-1).numpy()
            wrong = torch.cat([h[:, 0], h[:, 1:].mean(1)], dim=-1).numpy()          # what a head that keeps the registers would give
        ext = B200FeatureExtractor(name, sd, input_patch=P, max_batch=4)
        got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, torch.from_numpy(rows).cuda()).cpu().numpy()
        ext.cleanup()
        assert got.shape == want.shape == (n, 768)
        for half in (slice(0, 384), slice(384, 768)):
            r = np.linalg.norm(got[:, half] - want[:, half], axis=1) / np.linalg.norm(want[:, half], axis=1)
            assert r.max() < 1e-3, r
        assert (np.linalg.norm(wrong - want, axis=1) / np.linalg.norm(want, axis=1)).min() > 5e-3    # the test can tell the two apart
    finally:
        for d in (wt.DINOV2_SPECS, wt.DINOV2_REGISTERS, d2.DINOV2_CONFIGS, d2.DINOV2_REGISTERS, FAMILY_RECIPES):
            d.pop(name, None)
