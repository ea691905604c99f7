"""SAM2 tissue-mask predictor on the B200 kernels behind the reference's predictor contract.

`B200Sam2Predictor.predict_logits(uint8 (1024,1024,3)) -> float32 (1024,1024)` is what
`segmentation.B200SegmentationService` needs; it replaces SAM2ImagePredictor.set_image + predict(box=whole image)
(atlas_patch/services/segmentation.py:127-136).  Weights are given as a state_dict with transformers' Sam2Model names.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np

from atlaspatch_b200._lib import Context, Sam2Desc

# atlas_patch/configs/sam2.1_hiera_t.yaml:10-15 (the model the reference CLI ships, cli.py:46-47)
HIERA_T = dict(embed_dim=96, blocks=(1, 2, 7, 2), heads=(1, 2, 4, 8), windows=(8, 4, 14, 7), global_blocks=(5, 7, 9))


class B200Sam2Predictor:
    def __init__(self, state_dict: Mapping[str, object], *, config: dict = HIERA_T, device: int = 0):
        self.ctx = Context.get(device)
        lib = self.ctx.lib
        gb = tuple(config["global_blocks"])
        desc = Sam2Desc(embed_dim=config["embed_dim"], blocks_per_stage=(C.c_int * 4)(*config["blocks"]),
                        heads_per_stage=(C.c_int * 4)(*config["heads"]), window_per_stage=(C.c_int * 4)(*config["windows"]),
                        n_global=len(gb), global_blocks=(C.c_int * 8)(*(gb + (0,) * (8 - len(gb)))))
        h = C.c_void_p()
        self.ctx.check(lib.ap_sam2_create(self.ctx.handle, C.byref(desc), C.byref(h)))
        self._h = h
        try:
            for key, t in state_dict.items():
                a = t.detach().to("cpu").float().contiguous().numpy() if hasattr(t, "detach") else np.ascontiguousarray(t, np.float32)
                self.ctx.check(lib.ap_sam2_set_tensor(h, key.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            self.ctx.check(lib.ap_sam2_finalize(h))
        except Exception:
            lib.ap_sam2_destroy(h)
            self._h = None
            raise

    def predict_logits(self, image_u8: np.ndarray, return_lowres: bool = False):
        a = np.ascontiguousarray(image_u8)
        if a.shape != (1024, 1024, 3) or a.dtype != np.uint8:
            raise ValueError(f"expected uint8 (1024,1024,3), got {a.dtype} {a.shape}")
        logits = np.empty((1024, 1024), dtype=np.float32)
        low = np.empty((256, 256), dtype=np.float32)
        self.ctx.check(self.ctx.lib.ap_sam2_predict_host(self._h, a.ctypes.data_as(C.c_void_p), logits.ctypes.data_as(C.c_void_p),
                                                         low.ctypes.data_as(C.c_void_p)))
        return (logits, low) if return_lowres else logits

    def predict_logits_batch(self, images_u8) -> np.ndarray:
        """(n, 1024, 1024, 3) uint8 -> (n, 1024, 1024) float32 logits: the reference's predict_batch (services/segmentation.py:142-180)."""
        a = np.ascontiguousarray(np.stack([np.asarray(i) for i in images_u8]) if not isinstance(images_u8, np.ndarray) else images_u8)
        if a.ndim != 4 or a.shape[1:] != (1024, 1024, 3) or a.dtype != np.uint8:
            raise ValueError(f"expected uint8 (n,1024,1024,3), got {a.dtype} {a.shape}")
        logits = np.empty((a.shape[0], 1024, 1024), dtype=np.float32)
        if a.shape[0]:
            self.ctx.check(self.ctx.lib.ap_sam2_predict_batch_host(self._h, a.ctypes.data_as(C.c_void_p), int(a.shape[0]),
                                                                   logits.ctypes.data_as(C.c_void_p), None))
        return logits

    def debug_buffer(self, name: str, shape) -> np.ndarray:
        out = np.empty(shape, dtype=np.float32)
        self.ctx.check(self.ctx.lib.ap_sam2_debug_copy(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self.ctx.lib.ap_sam2_destroy(self._h)
            self._h = None

    __del__ = close


HIERA_L = dict(embed_dim=144, blocks=(2, 6, 36, 4), heads=(2, 4, 8, 16), windows=(8, 4, 16, 8), global_blocks=(23, 33, 43))

_UPSTREAM_RULES = [   # (upstream substring, HF substring); applied in order to every key of checkpoint["model"]
    ("image_encoder.trunk.patch_embed.proj.", "vision_encoder.backbone.patch_embed.projection."),
    ("image_encoder.trunk.", "vision_encoder.backbone."),
    ("image_encoder.neck.", "vision_encoder.neck."),
    (".conv.weight", ".weight"), (".conv.bias", ".bias"),                    # neck.convs.N.conv.* -> neck.convs.N.*
    (".norm1.", ".layer_norm1."), (".norm2.", ".layer_norm2."), (".norm3.", ".layer_norm3."), (".norm4.", ".layer_norm4."),
    ("sam_mask_decoder.", "mask_decoder."), ("sam_prompt_encoder.", "prompt_encoder."),
    (".out_proj.", ".o_proj."), ("norm_final_attn", "layer_norm_final_attn"),
    ("output_upscaling.0.", "upscale_conv1."), ("output_upscaling.1.", "upscale_layer_norm."), ("output_upscaling.3.", "upscale_conv2."),
    ("pe_layer.positional_encoding_gaussian_matrix", "shared_embedding.positional_embedding"),
    ("mask_downscaling.0.", "mask_embed.conv1."), ("mask_downscaling.1.", "mask_embed.layer_norm1."),
    ("mask_downscaling.3.", "mask_embed.conv2."), ("mask_downscaling.4.", "mask_embed.layer_norm2."),
    ("mask_downscaling.6.", "mask_embed.conv3."),
    ("no_mem_embed", "no_memory_embedding"),
]


def convert_upstream_state_dict(model_sd: Mapping[str, object]) -> dict:
    """facebookresearch/sam2 parameter names (what the reference loads: `torch.load(model.pth)["model"]`,
    atlas_patch/services/segmentation.py:66-67) -> transformers' Sam2Model names used by this engine.

    NOTE: written from the published module layout; neither the `sam2` package nor a checkpoint is available offline, so this
    map is exercised only by a round-trip test on synthetic keys.  Memory-attention / memory-encoder tensors (instantiated
    by the reference but never executed on the image path, configs/sam2.1_hiera_t.yaml:30-86) are dropped.
    """
    import numpy as np

    out: dict = {}
    points: dict[int, object] = {}
    for key, val in model_sd.items():
        if key.startswith(("memory_attention.", "memory_encoder.", "obj_ptr", "mask_downsample", "maskmem_tpos_enc",
                           "no_mem_pos_enc", "no_obj_ptr", "no_obj_embed_spatial")):
            continue
        k = key
        for a, b in _UPSTREAM_RULES:
            k = k.replace(a, b)
        if ".mlp.layers." in k or "_head.layers." in k or "hypernetworks_mlps" in k:
            # MLP(layers=[Linear...]) -> proj_in / layers.{i-1} / proj_out
            head, idx_rest = k.rsplit(".layers.", 1)
            idx, rest = idx_rest.split(".", 1)
            n_layers = 2 if ".mlp" in head and "transformer" in head or ".blocks." in head else 3
            i = int(idx)
            k = f"{head}.proj_in.{rest}" if i == 0 else (f"{head}.proj_out.{rest}" if i == n_layers - 1 else f"{head}.layers.{i - 1}.{rest}")
        if "prompt_encoder.point_embeddings." in k:
            points[int(k.split("point_embeddings.")[1].split(".")[0])] = val
            continue
        out[k] = val
    if points:
        to_np = lambda t: t.detach().cpu().float().numpy() if hasattr(t, "detach") else np.asarray(t, np.float32)
        out["prompt_encoder.point_embed.weight"] = np.concatenate([to_np(points[i]).reshape(1, -1) for i in sorted(points)], axis=0)
    g = out.get("prompt_encoder.shared_embedding.positional_embedding")
    if g is not None:
        out["shared_image_embedding.positional_embedding"] = g
    return out
