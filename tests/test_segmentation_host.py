"""CPU tests: the host-side segmentation steps (a2, a3, a5) against the reference's own functions when it is importable
(build container), and their invariants everywhere."""
import numpy as np
import pytest
from PIL import Image

from atlaspatch_b200 import segmentation as seg
from oracle.refimport import import_reference, reference_available


def test_mask_roundtrip_and_threshold():
    rng = np.random.default_rng(0)
    m = (rng.random((1024, 1024)) > 0.5).astype(np.float32)
    small = seg.resize_mask(m, (768, 1024))
    assert small.shape == (768, 1024) and set(np.unique(small)) <= {0.0, 1.0}
    img = rng.integers(0, 256, (600, 800, 3), dtype=np.uint8)
    out, orig = seg.resize_for_sam(img)
    assert out.shape == (1024, 1024, 3) and orig == (600, 800)
    same, orig2 = seg.resize_for_sam(out)
    assert same is out and orig2 == (1024, 1024)
    t = seg.cap_thumbnail(Image.fromarray(rng.integers(0, 256, (3750, 5000, 3), dtype=np.uint8)))
    assert t.size == (1024, 768)


@pytest.mark.skipif(not reference_available(), reason="reference source only exists in the build container")
def test_host_steps_match_reference_functions():
    import_reference()
    from atlas_patch.services.segmentation import _SAM2Predictor

    ref = _SAM2Predictor.__new__(_SAM2Predictor)   # helpers only; no model is built
    ref.input_size = 1024
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (768, 1000, 3), dtype=np.uint8)
    a, sa = seg.resize_for_sam(img)
    b, sb = ref._resize_input_for_sam(img)
    assert sa == sb and np.array_equal(a, b)
    m = (rng.random((1024, 1024)) > 0.3).astype(np.float32)
    for shape in [(768, 1000), (512, 512), (313, 1024)]:
        assert np.array_equal(seg.resize_mask(m, shape), ref._resize_mask(m, shape))


def test_service_contract_with_a_stub_predictor():
    class W:  # minimal IWSI-like object
        def get_thumbnail_at_power(self, *, power, interpolation):
            return Image.fromarray(np.full((300, 500, 3), 200, np.uint8))

    svc = seg.B200SegmentationService(lambda im: np.where(np.arange(1024)[None, :] < 512, 1.0, -1.0) * np.ones((1024, 1)))
    mask = svc.segment_thumbnail(W())
    assert mask.data.shape == (300, 500) == mask.source_shape and mask.data.dtype == np.float32
    assert mask.data[:, :240].all() and not mask.data[:, 260:].any()
    with pytest.raises(NotImplementedError):
        seg.B200SegmentationService().segment_thumbnail(W())


def test_upstream_checkpoint_key_map_round_trip():
    """convert_upstream_state_dict: synthetic facebookresearch/sam2-style names -> every tensor the engine asks for."""
    from atlaspatch_b200.sam2 import convert_upstream_state_dict

    hf_needed = {
        "no_memory_embedding", "shared_image_embedding.positional_embedding", "vision_encoder.backbone.pos_embed",
        "vision_encoder.backbone.patch_embed.projection.weight", "vision_encoder.backbone.blocks.3.layer_norm1.weight",
        "vision_encoder.backbone.blocks.3.attn.qkv.bias", "vision_encoder.backbone.blocks.3.mlp.proj_in.weight",
        "vision_encoder.backbone.blocks.3.mlp.proj_out.bias", "vision_encoder.backbone.blocks.1.proj.weight",
        "vision_encoder.neck.convs.2.weight", "prompt_encoder.point_embed.weight", "prompt_encoder.not_a_point_embed.weight",
        "prompt_encoder.no_mask_embed.weight", "prompt_encoder.shared_embedding.positional_embedding",
        "mask_decoder.transformer.layers.1.self_attn.o_proj.weight", "mask_decoder.transformer.layers.0.layer_norm4.bias",
        "mask_decoder.transformer.layers.0.mlp.proj_in.weight", "mask_decoder.transformer.layers.0.mlp.proj_out.weight",
        "mask_decoder.transformer.layer_norm_final_attn.weight", "mask_decoder.upscale_conv1.weight",
        "mask_decoder.upscale_layer_norm.bias", "mask_decoder.upscale_conv2.bias",
        "mask_decoder.output_hypernetworks_mlps.0.proj_in.weight", "mask_decoder.output_hypernetworks_mlps.0.layers.0.weight",
        "mask_decoder.output_hypernetworks_mlps.0.proj_out.bias", "mask_decoder.conv_s0.weight", "mask_decoder.iou_token.weight",
    }
    upstream = {
        "no_mem_embed": 1, "image_encoder.trunk.pos_embed": 1, "image_encoder.trunk.patch_embed.proj.weight": 1,
        "image_encoder.trunk.blocks.3.norm1.weight": 1, "image_encoder.trunk.blocks.3.attn.qkv.bias": 1,
        "image_encoder.trunk.blocks.3.mlp.layers.0.weight": 1, "image_encoder.trunk.blocks.3.mlp.layers.1.bias": 1,
        "image_encoder.trunk.blocks.1.proj.weight": 1, "image_encoder.neck.convs.2.conv.weight": 1,
        "sam_prompt_encoder.point_embeddings.0.weight": np.zeros((1, 4)), "sam_prompt_encoder.point_embeddings.1.weight": np.ones((1, 4)),
        "sam_prompt_encoder.point_embeddings.2.weight": np.ones((1, 4)) * 2, "sam_prompt_encoder.point_embeddings.3.weight": np.ones((1, 4)) * 3,
        "sam_prompt_encoder.not_a_point_embed.weight": 1, "sam_prompt_encoder.no_mask_embed.weight": 1,
        "sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix": 7,
        "sam_mask_decoder.transformer.layers.1.self_attn.out_proj.weight": 1, "sam_mask_decoder.transformer.layers.0.norm4.bias": 1,
        "sam_mask_decoder.transformer.layers.0.mlp.layers.0.weight": 1, "sam_mask_decoder.transformer.layers.0.mlp.layers.1.weight": 1,
        "sam_mask_decoder.transformer.norm_final_attn.weight": 1, "sam_mask_decoder.output_upscaling.0.weight": 1,
        "sam_mask_decoder.output_upscaling.1.bias": 1, "sam_mask_decoder.output_upscaling.3.bias": 1,
        "sam_mask_decoder.output_hypernetworks_mlps.0.layers.0.weight": 1, "sam_mask_decoder.output_hypernetworks_mlps.0.layers.1.weight": 1,
        "sam_mask_decoder.output_hypernetworks_mlps.0.layers.2.bias": 1, "sam_mask_decoder.conv_s0.weight": 1,
        "sam_mask_decoder.iou_token.weight": 1, "memory_attention.layers.0.norm1.weight": 1, "maskmem_tpos_enc": 1,
    }
    out = convert_upstream_state_dict(upstream)
    assert hf_needed <= set(out), sorted(hf_needed - set(out))
    assert out["prompt_encoder.point_embed.weight"].shape == (4, 4) and out["prompt_encoder.point_embed.weight"][3, 0] == 3
    assert out["shared_image_embedding.positional_embedding"] == 7
    assert not any(k.startswith("memory_") for k in out)
