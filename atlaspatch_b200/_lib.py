"""ctypes binding of libatlaspatch_b200.so (the C ABI in include/atlaspatch_b200.h).

There is no CPU fallback: if the shared library is missing or cannot initialise a B200, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libatlaspatch_b200.so"

AP_OK, AP_EINVAL, AP_ECUDA, AP_ENOMEM, AP_ECAPACITY, AP_ESTATE = 0, -1, -2, -3, -4, -5
EPI_BIAS_F16, EPI_BIAS_GELU_F16, EPI_BIAS_RESID_F32, EPI_BIAS_F32 = 0, 1, 2, 3


class VitDesc(C.Structure):
    _fields_ = [("image_size", C.c_int), ("patch", C.c_int), ("layers", C.c_int), ("heads", C.c_int),
                ("hidden", C.c_int), ("mlp", C.c_int), ("input_patch", C.c_int), ("max_batch", C.c_int), ("precise_layers", C.c_int),
                ("ln_eps", C.c_float), ("mean", C.c_float * 3), ("std", C.c_float * 3),
                ("preprocess", C.c_int), ("resize_to", C.c_int), ("mlp_kind", C.c_int), ("pool", C.c_int), ("registers", C.c_int),
                ("pre_ln", C.c_int), ("proj_dim", C.c_int)]


class Sam2Desc(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("blocks_per_stage", C.c_int * 4), ("heads_per_stage", C.c_int * 4),
                ("window_per_stage", C.c_int * 4), ("n_global", C.c_int), ("global_blocks", C.c_int * 8)]


_P = C.c_void_p
_I32P = C.POINTER(C.c_int32)
_SIGNATURES = {
    "ap_version": (C.c_int, []),
    "ap_sizeof": (C.c_int, [C.c_char_p]),
    "ap_init": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ap_destroy": (C.c_int, [_P]),
    "ap_last_error": (C.c_char_p, [_P]),
    "ap_launch_count": (C.c_int64, [_P]),
    "ap_sm_count": (C.c_int, [_P]),
    "ap_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ap_profile_enable": (C.c_int, [_P, C.c_int]),
    "ap_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "ap_profile_read_tagged": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int,
                                         C.POINTER(C.c_int)]),
    "ap_synth_render": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, _P, C.c_int, _P, C.c_int,
                                  C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    "ap_thumbnail_area": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, _P, _P]),
    "ap_thumbnail_resize": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "ap_coords_capacity": (C.c_int64, [_P, _P, C.c_int, C.c_int]),
    "ap_extract_coords": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _P, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "ap_filter_patches": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_double, _P, _P, C.POINTER(C.c_int64), _P, _P]),
    "ap_encoder_create": (C.c_int, [_P, C.POINTER(VitDesc), C.POINTER(_P)]),
    "ap_encoder_destroy": (C.c_int, [_P]),
    "ap_encoder_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "ap_encoder_finalize": (C.c_int, [_P]),
    "ap_encoder_embedding_dim": (C.c_int, [_P]),
    "ap_linear_tap_tables": (C.c_int, [C.c_int, C.c_int, _I32P, C.POINTER(C.c_int16)]),
    "ap_resize_tap_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _I32P, _I32P, _I32P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ap_encoder_embed_coords": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int, _P, _P]),
    "ap_encoder_preprocess": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int, _P, C.POINTER(C.c_int64), _P]),
    "ap_encoder_embed_patches_host": (C.c_int, [_P, C.POINTER(_P), C.c_int64, _P]),
    "ap_sam2_create": (C.c_int, [_P, C.POINTER(Sam2Desc), C.POINTER(_P)]),
    "ap_sam2_destroy": (C.c_int, [_P]),
    "ap_sam2_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "ap_sam2_finalize": (C.c_int, [_P]),
    "ap_sam2_forward": (C.c_int, [_P, _P, _P, _P, _P]),
    "ap_sam2_predict_host": (C.c_int, [_P, _P, _P, _P]),
    "ap_sam2_predict_batch_host": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ap_sam2_debug_copy": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "ap_gemm_f16": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ap_gemm_f16_split": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ap_layernorm_f16": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_float, _P, C.c_int, C.c_int, _P]),
    "ap_attention_f16": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_lock = threading.Lock()


class AtlasB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"atlaspatch_b200 error {code}: {msg}")
        self.code = code


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library and bind every symbol the header declares."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(atlaspatch_b200 has no CPU fallback)")
        lib = C.CDLL(os.fspath(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        for name, struct in (("ap_vit_desc", VitDesc), ("ap_sam2_desc", Sam2Desc)):
            if lib.ap_sizeof(name.encode()) != C.sizeof(struct):
                raise ImportError(f"{LIB_PATH}: sizeof({name}) = {lib.ap_sizeof(name.encode())} but the ctypes declaration has "
                                  f"{C.sizeof(struct)} bytes (stale library? rebuild with __graft_entry__.build())")
        _lib = lib
        return lib


class Context:
    """One per process/GPU (ap_init).  Raises AtlasB200Error when there is no B200."""

    _instances: dict[int, "Context"] = {}

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.device = device
        h = _P()
        rc = self.lib.ap_init(device, C.byref(h))
        if rc != AP_OK:
            raise AtlasB200Error(rc, self.lib.ap_last_error(None).decode())
        self.handle = h

    @classmethod
    def get(cls, device: int = 0) -> "Context":
        if device not in cls._instances:
            cls._instances[device] = Context(device)
        return cls._instances[device]

    def check(self, rc: int) -> None:
        if rc != AP_OK:
            raise AtlasB200Error(rc, self.lib.ap_last_error(self.handle).decode())

    @property
    def launch_count(self) -> int:
        return int(self.lib.ap_launch_count(self.handle))

    @property
    def sm_count(self) -> int:
        return int(self.lib.ap_sm_count(self.handle))

    KERNEL_CLASSES = ("gemm", "attention", "layernorm", "preprocess", "coords", "thumbnail", "other")

    def set_option(self, key: str, value: int) -> None:
        self.check(self.lib.ap_set_option(self.handle, key.encode(), int(value)))

    def profile(self, on, classes=None) -> None:
        """on=False: off; on=True: time every kernel class, or only `classes` (names from KERNEL_CLASSES)."""
        mask = 0
        if on:
            mask = -1 if classes is None else sum(1 << self.KERNEL_CLASSES.index(c) for c in classes)
        self.check(self.lib.ap_profile_enable(self.handle, mask))

    def profile_read(self) -> dict[str, tuple[float, int]]:
        """{kernel class: (total ms, launches)} measured with CUDA events since the last read."""
        n = len(self.KERNEL_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self.check(self.lib.ap_profile_read(self.handle, ms, cnt, n))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_CLASSES)}


    def profile_read_gemm_shapes(self) -> dict[tuple[int, int, int], tuple[float, int]]:
        """{(N, K, epilogue): (total ms, launches)} of the GEMM launches timed since the last read (clears all records)."""
        cap = 64
        keys, ms, cnt, n = (C.c_int64 * cap)(), (C.c_double * cap)(), (C.c_int64 * cap)(), C.c_int(0)
        self.check(self.lib.ap_profile_read_tagged(self.handle, 0, keys, ms, cnt, cap, C.byref(n)))
        return {(int(keys[i]) >> 32, (int(keys[i]) & 0xFFFFFFFF) >> 4, int(keys[i]) & 15): (float(ms[i]), int(cnt[i])) for i in range(n.value)}


def current_stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)
