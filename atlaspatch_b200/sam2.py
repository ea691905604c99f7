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

    def debug_buffer(self, name: str, shape) -> np.ndarray:
        out = np.empty(shape, dtype=np.float32)
        self.ctx.check(self.ctx.lib.ap_sam2_debug_copy(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self.ctx.lib.ap_sam2_destroy(self._h)
            self._h = None

    __del__ = close
