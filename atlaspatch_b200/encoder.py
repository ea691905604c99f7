"""Encoder plug-in (seam #2 of SURVEY.md section 8b): the reference's FeatureExtractor contract on the B200 engine.

reference interface: atlas_patch/models/patch/base.py:15-29 (name, embedding_dim, extract_batch, cleanup);
extract_batch(patches: Sequence[np.ndarray (P,P,3) uint8], *, batch_size) -> np.ndarray (len, D) float32, rows in
input order, (0, D) for empty input (base.py:76-107).  `embed_coords` is the zero-copy fast path that reads patches
straight out of a device-resident slide.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from atlaspatch_b200._lib import Context, VitDesc, current_stream_ptr
from atlaspatch_b200.dinov2 import (DINOV2_CONFIGS, DINOV2_REGISTERS, HF_CLIP_CONFIGS, HF_VIT_CONFIGS, convert_dinov2_state_dict,
                                    convert_hf_clip_state_dict, convert_hf_vit_state_dict)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

# name -> (patch, layers, heads, hidden, mlp); the torchvision ViTs the reference registers (models/patch/vit.py:9-15)
VIT_CONFIGS = {
    "vit_b_16": (16, 12, 12, 768, 3072),
    "vit_b_32": (32, 12, 12, 768, 3072),
    "vit_l_16": (16, 24, 16, 1024, 4096),
    "vit_l_32": (32, 24, 16, 1024, 4096),
    # vit_h_14 (models/patch/vit.py:14) has head_dim 80: not built -- the attention kernels are head_dim 64
    "vit_test_tiny": (16, 2, 4, 256, 512),
}

# Hub encoders of the reference that run on the same kernels with their own preprocess / head.  The
# processor settings are the published contents of each repo's preprocessor_config.json (no network here: restated, not fetched).
#   preprocess: ap_vit_desc.preprocess (1 ATen uint8 bicubic-antialias, 2 Pillow BILINEAR, 3 ATen uint8 bilinear-antialias, 4 Pillow BICUBIC)
#   pool: 0 class token, 1 [class || mean of patch tokens]
_HALF = (0.5, 0.5, 0.5)
_HIBOU_MEAN, _HIBOU_STD = (0.7068, 0.5755, 0.722), (0.195, 0.2316, 0.1816)
_CLIP_MEAN, _CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
_HOPT = dict(preprocess=2, resize_to=224, mean=(0.707223, 0.578729, 0.703617), std=(0.211883, 0.230117, 0.177517), pool=0, ln_eps=1e-6,
             default_patch=224)
_OCLIP = dict(preprocess=4, resize_to=224, mean=_CLIP_MEAN, std=_CLIP_STD, pool=0, ln_eps=1e-5, default_patch=224)
_PORCH = dict(preprocess=2, resize_to=224, pool=0, ln_eps=1e-6, default_patch=224)
_GIGA = dict(preprocess=4, resize_to=256, pool=0, ln_eps=1e-6, default_patch=256)
_CLIP = dict(preprocess=1, resize_to=224, mean=_CLIP_MEAN, std=_CLIP_STD, pool=0, ln_eps=1e-5, default_patch=224)
FAMILY_RECIPES = {
    # kaiko-ai/midnight (models/patch/midnight.py:15-25,55-61): torchvision Resize(224) on the PIL patch, CenterCrop(224),
    # Normalize(0.5, 0.5); feature = cat(last_hidden_state[:, 0], last_hidden_state[:, 1:].mean(1)) -> 3072
    "midnight": dict(preprocess=2, resize_to=224, mean=_HALF, std=_HALF, pool=1, ln_eps=1e-6, default_patch=224),
    "midnight_test_tiny": dict(preprocess=2, resize_to=224, mean=_HALF, std=_HALF, pool=1, ln_eps=1e-6, default_patch=224),
    # owkin/phikon-v2 (models/patch/phikon.py:90-93,103-105): BitImageProcessor(fast) shortest_edge 224 bicubic, crop 224, ImageNet
    "phikon_v2": dict(preprocess=1, resize_to=224, pool=0, ln_eps=1e-6, default_patch=224),
    "phikon_v2_test_tiny": dict(preprocess=1, resize_to=224, pool=0, ln_eps=1e-6, default_patch=224),
    # owkin/phikon (models/patch/phikon.py:41-46,54-56): ViTImageProcessor(fast) 224 x 224 resample 2 (bilinear), ImageNet; ViTModel
    # with layer_norm_eps 1e-12
    "phikon_v1": dict(preprocess=3, resize_to=224, pool=0, ln_eps=1e-12, default_patch=224),
    "phikon_v1_test_tiny": dict(preprocess=3, resize_to=224, pool=0, ln_eps=1e-12, default_patch=224),
    # histai/hibou-B / -L (models/patch/hibou.py:51-54,67-69): BitImageProcessor(fast) shortest_edge 224 bicubic, crop 224, the
    # repo's own mean / std; DINOv2 with 4 register tokens, feature = pooler_output (class token after the final LayerNorm)
    "hibou_b": dict(preprocess=1, resize_to=224, mean=_HIBOU_MEAN, std=_HIBOU_STD, pool=0, ln_eps=1e-6, default_patch=224),
    "hibou_l": dict(preprocess=1, resize_to=224, mean=_HIBOU_MEAN, std=_HIBOU_STD, pool=0, ln_eps=1e-6, default_patch=224),
    "hibou_test_tiny": dict(preprocess=1, resize_to=224, mean=_HIBOU_MEAN, std=_HIBOU_STD, pool=0, ln_eps=1e-6, default_patch=224),
    # SophontAI/OpenMidnight (models/patch/openmidnight.py:17-30,49): torchvision Resize((224, 224)) on the PIL patch, ImageNet
    # mean / std; facebookresearch dinov2_vitg14_reg (4 register tokens, SwiGLU), feature = class token
    "openmidnight": dict(preprocess=2, resize_to=224, pool=0, ln_eps=1e-6, default_patch=224),
    "openmidnight_test_tiny": dict(preprocess=2, resize_to=224, pool=0, ln_eps=1e-6, default_patch=224),
    # bioptimus/H-optimus-0 / -1 (models/patch/hoptimus.py:14-31): torchvision Resize((224, 224)) on the PIL patch, the models' own mean / std
    "h_optimus_0": _HOPT, "h_optimus_1": _HOPT, "h_optimus_test_tiny": _HOPT,
    # AI4Pathology/PathOrchestra (models/patch/pathorchestra.py:52-58): torchvision Resize(224) on the PIL patch, ImageNet mean / std
    "pathorchestra": _PORCH, "pathorchestra_test_tiny": _PORCH,
    # prov-gigapath (models/patch/gigapath.py:17-26): torchvision Resize(256, BICUBIC) -> CenterCrop(224) on the PIL patch, ImageNet
    "prov_gigapath": _GIGA, "prov_gigapath_test_tiny": _GIGA,
    # vinid/plip, wisdomik/QuiltNet-B-32 / -B-16 (models/patch/plip.py:34-35,56, quilt.py:56-60): transformers CLIPModel +
    # CLIPProcessor (fast image processor: shortest_edge 224 bicubic, crop 224, OpenAI CLIP mean / std), feature =
    # get_image_features = visual_projection(post_layernorm(class token)) -> 512
    "plip": _CLIP, "quilt_b_32": _CLIP, "quilt_b_16": _CLIP, "plip_test_tiny": _CLIP, "quilt_b_16_test_tiny": _CLIP,
    # OpenAI CLIP through open_clip (models/patch/clip.py:36-40,58): open_clip's eval transform = torchvision Resize(224, BICUBIC) ->
    # CenterCrop(224) on the PIL patch (Pillow BICUBIC), CLIP mean / std; feature = encode_image = ln_post(class token) @ visual.proj
    "clip_vit_b_32": _OCLIP, "clip_vit_b_16": _OCLIP, "clip_vit_l_14": _OCLIP, "clip_vit_b_32_test_tiny": _OCLIP,
    "clip_vit_l_14_test_tiny": _OCLIP,
}


def vit_state_dict_names(layers: int, registers: int = 0, pre_ln: bool = False, proj: bool = False) -> list[str]:
    names = ["conv_proj.weight", "conv_proj.bias", "class_token", "encoder.pos_embedding", "encoder.ln.weight", "encoder.ln.bias"]
    if registers:
        names.append("register_tokens")
    if pre_ln:
        names += ["encoder.pre_ln.weight", "encoder.pre_ln.bias"]
    if proj:
        names.append("head.proj.weight")
    for i in range(layers):
        p = f"encoder.layers.encoder_layer_{i}."
        names += [p + s for s in ("ln_1.weight", "ln_1.bias", "self_attention.in_proj_weight", "self_attention.in_proj_bias",
                                  "self_attention.out_proj.weight", "self_attention.out_proj.bias", "ln_2.weight", "ln_2.bias",
                                  "mlp.0.weight", "mlp.0.bias", "mlp.3.weight", "mlp.3.bias")]
    return names


class B200FeatureExtractor:
    """ViT encoder forward on hand-written sm_100a kernels behind the reference's FeatureExtractor interface.

    torchvision ViTs (models/patch/vit.py) take a torchvision state_dict and the ImageClassification preset (centre crop);
    `dinov2_*` (models/patch/dinov2.py) take a transformers Dinov2Model state_dict and the BitImageProcessorFast preprocess
    (bicubic-antialias resize to 256, centre crop 224).  The hub families of FAMILY_RECIPES (midnight, phikon, hibou, openmidnight,
    h_optimus, pathorchestra, prov_gigapath, plip / quilt, OpenAI CLIP) take the state_dict of the model class the reference loads, in that
    class's own key layout (transformers, facebookresearch dinov2, timm, open_clip), and run their preprocess, register tokens and head
    ([class || mean], visual projection) in the CUDA path.  `input_patch` is the --patch-size of the run (the size of the patches
    handed to extract_batch / cut by embed_coords)."""

    def __init__(self, name: str, state_dict: Mapping[str, "object"], *, input_patch: int | None = None, image_size: int = 224,
                 max_batch: int = 254, device: int = 0, config: tuple | None = None, registry_name: str | None = None,
                 precise_layers: int = -1, precision: str = "fast", arch: tuple | None = None, recipe: dict | None = None):
        """max_batch: patches per forward chunk = size of the activation workspaces (ViT-B/16: ~1 GB at 254).  Larger chunks amortise
        launch gaps and the tails of the HBM-bound residual GEMMs: 127 -> 254 -> 508 patches gave 22.9 -> 24.1 -> 24.9 k patches/s for
        ViT-B/16 (bench.py uses 508), +2-5 % for ViT-L / DINOv2.
        precision: "fast" = the library defaults (typical rows 5.5e-4 of the fp32 path, p99 9.3e-4; rows with
        two large flat regions -- a patch hanging 40 % over the slide edge next to white background -- touch 1.0e-3);
        "strict" = two leading layers with split weights AND split A operands (they run their LayerNorm kernels, the other layers stay
        folded): max 8.7e-4 on the same 1 024-row survey (profiles/r02_vit_b_16_precision_survey.log), 21.1 k against 24.1 k patches/s."""
        if precision not in ("fast", "strict"):
            raise ValueError("precision must be 'fast' or 'strict'")
        if (arch is None) != (recipe is None):
            raise ValueError("arch and recipe are given together")
        preprocess, resize_to, mlp_kind, pool, ln_eps, registers, pre_ln, proj_dim = 0, 0, 0, 0, 1e-6, 0, 0, 0
        mean, std = IMAGENET_MEAN, IMAGENET_STD
        if recipe is None:
            recipe = FAMILY_RECIPES.get(name, {}) if config is None else {}
        if config is None and (arch is not None or name in DINOV2_CONFIGS or name in HF_VIT_CONFIGS or name in HF_CLIP_CONFIGS):
            if arch is not None:
                # a DINOv2-layout ViT (pre-LayerNorm blocks, optional LayerScale / SwiGLU / register tokens) described by the caller:
                # (patch, layers, heads, hidden, mlp, swiglu, registers) + a FAMILY_RECIPES-style recipe.  plugin.py derives both from a
                # timm model object and its data config for the timm-loaded families (uni, h0_mini, lunit)
                patch, layers, heads, hidden, mlp, swiglu, registers = arch
                state_dict = convert_dinov2_state_dict(state_dict, layers=layers, swiglu=swiglu, image_size=image_size, patch=patch,
                                                       registers=registers)
            elif name in HF_CLIP_CONFIGS:
                (patch, layers, heads, hidden, mlp, proj_dim), swiglu, pre_ln = HF_CLIP_CONFIGS[name], False, 1
                state_dict = convert_hf_clip_state_dict(state_dict, layers=layers)
            elif name in DINOV2_CONFIGS:
                patch, layers, heads, hidden, mlp, swiglu = DINOV2_CONFIGS[name]
                registers = DINOV2_REGISTERS.get(name, 0)
                state_dict = convert_dinov2_state_dict(state_dict, layers=layers, swiglu=swiglu, image_size=image_size, patch=patch,
                                                       registers=registers)
            else:
                (patch, layers, heads, hidden, mlp), swiglu = HF_VIT_CONFIGS[name], False
                state_dict = convert_hf_vit_state_dict(state_dict, layers=layers)
            preprocess, resize_to, mlp_kind = recipe.get("preprocess", 1), recipe.get("resize_to", 256), (2 if name in HF_CLIP_CONFIGS else int(swiglu))
            pool, ln_eps = recipe.get("pool", 0), recipe.get("ln_eps", 1e-6)
            mean, std = recipe.get("mean", IMAGENET_MEAN), recipe.get("std", IMAGENET_STD)
            input_patch = recipe.get("default_patch", 224) if input_patch is None else input_patch
        else:
            cfg = config or VIT_CONFIGS.get(name)
            if cfg is None:
                raise KeyError(f"Unknown B200 encoder '{name}'. Available: {sorted(VIT_CONFIGS) + sorted(DINOV2_CONFIGS) + sorted(HF_VIT_CONFIGS) + sorted(HF_CLIP_CONFIGS)}")
            patch, layers, heads, hidden, mlp = cfg
            input_patch = 256 if input_patch is None else input_patch
            if int(input_patch) != 256:
                # --patch-size other than the preset's resize_size: torchvision resizes the PIL patch to 256 with Pillow's BILINEAR
                # before the 224 crop (models/patch/base.py:170 -> [tv]transforms/_presets.py:58-64); done in the CUDA preprocess
                preprocess, resize_to = 2, 256
        self._patch, self._grid = int(patch), int(image_size) // int(patch)
        self.name = registry_name or name   # H5 dataset name: features/<name> (services/storage.py:250-337)
        self.embedding_dim = int(proj_dim) if proj_dim else int(hidden) * (2 if pool == 1 else 1)
        self._mean = tuple(float(m) for m in mean)
        # the device path resamples reads larger than the patch (40x slide at 20x: cv2.resize, feature_embedding.py:93-95) only in front of
        # the crop preprocess; the resizing preprocesses take reads of exactly input_patch (services.py falls back to host reads otherwise)
        self.supports_large_reads = preprocess == 0
        self.input_patch = int(input_patch)
        self.max_batch = int(max_batch)
        self.ctx = Context.get(device)
        lib = self.ctx.lib
        if precision == "strict" and precise_layers < 2:
            precise_layers = 2 if layers <= 32 else 12
        desc = VitDesc(image_size=image_size, patch=patch, layers=layers, heads=heads, hidden=hidden, mlp=mlp,
                       input_patch=input_patch, max_batch=max_batch, precise_layers=precise_layers, ln_eps=ln_eps,
                       mean=(C.c_float * 3)(*mean), std=(C.c_float * 3)(*std), preprocess=preprocess,
                       resize_to=resize_to, mlp_kind=mlp_kind, pool=pool, registers=registers, pre_ln=pre_ln, proj_dim=proj_dim)
        h = C.c_void_p()
        self.ctx.check(lib.ap_encoder_create(self.ctx.handle, C.byref(desc), C.byref(h)))
        self._h = h
        try:
            for key in vit_state_dict_names(layers, registers, bool(pre_ln), bool(proj_dim)):
                if key not in state_dict:
                    raise KeyError(f"state_dict is missing '{key}'")
                t = state_dict[key]
                a = t.detach().to("cpu").float().contiguous().numpy() if hasattr(t, "detach") else np.ascontiguousarray(t, np.float32)
                self.ctx.check(lib.ap_encoder_set_tensor(h, key.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            if precision == "strict":     # read at finalize; context-wide, so put the default back.  The layers whose A operands are
                self.ctx.set_option("precise_aw_layers", 2 if layers <= 32 else 12)   # split run their LayerNorm kernels, the rest stay folded
            try:
                self.ctx.check(lib.ap_encoder_finalize(h))
            finally:
                if precision == "strict":
                    self.ctx.set_option("precise_aw_layers", -1)
        except Exception:
            lib.ap_encoder_destroy(h)
            self._h = None
            raise

    # ---- reference contract ------------------------------------------------------------------------
    def extract_batch(self, patches: Sequence[np.ndarray], *, batch_size: int | None = None) -> np.ndarray:
        n = len(patches)
        out = np.empty((n, self.embedding_dim), dtype=np.float32)
        if n == 0:
            return out
        P = self.input_patch
        keep = []  # keep converted arrays alive while the C call runs
        ptrs = (C.c_void_p * n)()
        for i, p in enumerate(patches):
            a = np.asarray(p)
            if a.shape != (P, P, 3) or a.dtype != np.uint8:
                raise ValueError(f"patch {i}: expected uint8 array of shape ({P},{P},3), got {a.dtype} {a.shape}")
            if not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)
            keep.append(a)
            ptrs[i] = a.ctypes.data
        self.ctx.check(self.ctx.lib.ap_encoder_embed_patches_host(self._h, ptrs, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def cleanup(self) -> None:
        if getattr(self, "_h", None) is not None:
            self.ctx.lib.ap_encoder_destroy(self._h)
            self._h = None

    __del__ = cleanup

    def preprocess_pixels(self, image, W: int, H: int, pitch: int, coords_dev, read_size: int | None = None) -> np.ndarray:
        """a12 alone: the uint8 pixels (n, image, image, 3) the encoder sees after crop / resize, decoded from the im2col rows
        (parity tests; n <= max_batch)."""
        import torch

        n = int(coords_dev.shape[0])
        g, p = self._grid, self._patch
        kp = (3 * p * p + 63) // 64 * 64
        buf = torch.empty((max(n, 1) * g * g, kp), dtype=torch.float16, device="cuda")
        cols = C.c_int64(0)
        self.ctx.check(self.ctx.lib.ap_encoder_preprocess(
            self._h, C.c_void_p(image.data_ptr()), W, H, pitch, C.c_void_p(coords_dev.contiguous().data_ptr()), n,
            int(read_size or self.input_patch), C.c_void_p(buf.data_ptr()), C.byref(cols), C.c_void_p(current_stream_ptr())))
        assert cols.value == kp
        a = buf[:n * g * g, :3 * p * p].float().cpu().numpy() * 256.0           # (n g g, 3 p p) = pixel - centre
        a = a.reshape(n, g, g, 3, p, p).transpose(0, 1, 4, 2, 5, 3).reshape(n, g * p, g * p, 3)
        centre = np.rint(255.0 * np.asarray(self._mean, dtype=np.float64)).astype(np.float32)   # encoder.cu: lrint(255 * mean_c)
        return np.rint(a + centre).astype(np.uint8)

    # ---- device-resident fast path -----------------------------------------------------------------
    def embed_coords(self, image, W: int, H: int, pitch: int, coords_dev, out=None, read_size: int | None = None):
        """image: uint8 CUDA tensor (level-0 RGB rows, `pitch` bytes/row); coords_dev: int32 CUDA (n,5) -> (n,D) fp32 CUDA.
        read_size: the rows' read_w (= read_h); defaults to the patch size (no resize); any integer multiple of it is supported."""
        import torch

        n = int(coords_dev.shape[0])
        if out is None:
            out = torch.empty((n, self.embedding_dim), dtype=torch.float32, device="cuda")
        if n == 0:
            return out
        assert coords_dev.dtype == torch.int32 and coords_dev.is_contiguous() and coords_dev.is_cuda
        self.ctx.check(self.ctx.lib.ap_encoder_embed_coords(
            self._h, C.c_void_p(image.data_ptr()), W, H, pitch, C.c_void_p(coords_dev.data_ptr()), n,
            int(read_size or self.input_patch), C.c_void_p(out.data_ptr()), C.c_void_p(current_stream_ptr())))
        return out
