"""Device-resident slides: the synthetic backend (seam #3 of SURVEY.md section 8b) and the thumbnail kernel wrapper.

`SyntheticWSI` exposes the attributes / methods of the reference's IWSI (atlas_patch/core/wsi/iwsi.py:9-124:
w, h, nlvl, ds, dims, mpp, mag, extract, get_size, get_thumbnail_at_power, cleanup) so the adapters -- and the
reference's own services, when it is importable -- can use it, and additionally hands the HBM pointer of the
level-0 image to the kernels (`device_image`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from atlaspatch_b200._lib import Context, current_stream_ptr
from atlaspatch_b200.geometry import infer_mag, optimal_level, thumbnail_factor, validate_mpp
from atlaspatch_b200.synthetic import SyntheticSlideSpec


def render_device(spec: SyntheticSlideSpec, x: int, y: int, w: int, h: int, *, ctx: Context | None = None):
    """Region of the synthetic slide rendered by the CUDA generator -> torch uint8 (h, w, 3) on the GPU."""
    import torch

    ctx = ctx or Context.get(torch.cuda.current_device())
    pitch = (w * 3 + 15) // 16 * 16
    buf = torch.empty((h, pitch), dtype=torch.uint8, device="cuda")
    blobs = np.ascontiguousarray(spec.blob_array())
    holes = np.ascontiguousarray(spec.hole_array())
    ctx.check(ctx.lib.ap_synth_render(
        ctx.handle, C.c_void_p(buf.data_ptr()), pitch, spec.width, spec.height, spec.seed & 0xFFFFFFFF,
        blobs.ctypes.data_as(C.c_void_p), blobs.shape[0], holes.ctypes.data_as(C.c_void_p), holes.shape[0],
        x, y, w, h, C.c_void_p(current_stream_ptr())))
    return buf, pitch


def thumbnail_area(image, W: int, H: int, pitch: int, factor: int, *, ctx: Context | None = None):
    """a1 on the device: (H/f, W/f, 3) uint8 CUDA tensor = cv2.resize(INTER_AREA) of the level-0 image."""
    import torch

    ctx = ctx or Context.get(torch.cuda.current_device())
    if W % factor or H % factor:
        raise ValueError(f"thumbnail factor {factor} must divide the level size {W}x{H}")
    out = torch.empty((H // factor, W // factor, 3), dtype=torch.uint8, device="cuda")
    ctx.check(ctx.lib.ap_thumbnail_area(ctx.handle, C.c_void_p(image.data_ptr()), W, H, pitch, factor,
                                        C.c_void_p(out.data_ptr()), C.c_void_p(current_stream_ptr())))
    return out


def thumbnail_resize(image, W: int, H: int, pitch: int, out_w: int, out_h: int, *, ctx: Context | None = None):
    """a1, any size: (out_h, out_w, 3) uint8 CUDA tensor = cv2.resize(level, (out_w, out_h), INTER_AREA), bit for bit, for a
    down-scale (integer-factor kernels when both scale factors are integral, OpenCV's fractional cell weights otherwise)."""
    import torch

    ctx = ctx or Context.get(torch.cuda.current_device())
    out = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device="cuda")
    ctx.check(ctx.lib.ap_thumbnail_resize(ctx.handle, C.c_void_p(image.data_ptr()), W, H, pitch, int(out_w), int(out_h),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(current_stream_ptr())))
    return out


class DeviceWSI:
    """Single-level slide whose level-0 RGB image lives in HBM: the IWSI attribute / method contract of the reference
    (core/wsi/iwsi.py:9-124) plus `device_image` / `pitch` for the kernels.  Subclasses provide `_load_image()`."""

    def __init__(self, path: str, width: int, height: int, mpp: float, *, ctx: Context | None = None):
        self.path = path
        self.w, self.h = int(width), int(height)
        self.nlvl, self.ds, self.dims = 1, [1.0], [(self.w, self.h)]
        self.meta: dict = {}
        self.mpp = validate_mpp(float(mpp), source="manual")
        self.mag = infer_mag(self.mpp)
        self._ctx = ctx
        self._image = None
        self._pitch = 0

    def _load_image(self):
        raise NotImplementedError

    # ---- device side ----
    @property
    def device_image(self):
        """(H, pitch) uint8 CUDA tensor holding RGB HWC rows; produced on first use."""
        if self._image is None:
            self._image, self._pitch = self._load_image()
        return self._image

    @property
    def pitch(self) -> int:
        self.device_image
        return self._pitch

    # ---- IWSI contract ----
    def get_size(self, lv: int = 0) -> tuple[int, int]:
        return self.w, self.h

    def optimal_level(self, target_ds: float) -> tuple[int, float]:
        return optimal_level(self.ds, target_ds)

    def extract(self, xy, lv, wh, *, mode: str = "array"):
        """(h, w, 3) uint8 RGB numpy array read back from HBM; pixels outside the slide are 0."""
        import torch

        x, y = int(xy[0]), int(xy[1])
        w, h = int(wh[0]), int(wh[1])
        out = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + w, self.w), min(y + h, self.h)
        if x1 > x0 and y1 > y0:
            img = self.device_image
            out[y0 - y:y1 - y, x0 - x:x1 - x] = img[y0:y1, x0 * 3:x1 * 3].reshape(y1 - y0, x1 - x0, 3)
        arr = out.cpu().numpy()
        if mode == "image":
            from PIL import Image

            return Image.fromarray(arr)
        return arr

    def thumbnail_at_power_device(self, power: float = 1.25):
        """iwsi.py:246-323 on the device (single level): ds = mag / power, output round(W / ds) x round(H / ds) (Python round),
        cv2.resize(INTER_AREA) semantics for any level size."""
        ds = thumbnail_factor(self.mag, power)
        out_w, out_h = max(1, int(round(self.w / ds))), max(1, int(round(self.h / ds)))
        if (out_w, out_h) == (self.w, self.h):
            return self.device_image[:, :self.w * 3].reshape(self.h, self.w, 3).clone()
        if out_w > self.w or out_h > self.h:
            raise NotImplementedError(f"thumbnail power {power} above the slide's magnification {self.mag} (cubic up-scaling) is not built")
        return thumbnail_resize(self.device_image, self.w, self.h, self.pitch, out_w, out_h, ctx=self._ctx)

    def get_thumbnail_at_power(self, *, power: float = 1.25, interpolation: str = "optimise"):
        from PIL import Image

        return Image.fromarray(self.thumbnail_at_power_device(power).cpu().numpy())

    def get_thumb(self, max_hw):
        t = self.get_thumbnail_at_power(power=1.25)
        t.thumbnail(max_hw)
        return t

    def metadata_attrs(self) -> dict:
        return {"mpp": self.mpp, "magnification": int(self.mag)}

    def cleanup(self) -> None:
        self._image = None


class SyntheticWSI(DeviceWSI):
    """Synthetic slide generated on the device (never materialised on the host)."""

    def __init__(self, spec: SyntheticSlideSpec, *, path: str | None = None, ctx: Context | None = None):
        super().__init__(path or f"synthetic_{spec.width}x{spec.height}_s{spec.seed}.synth", spec.width, spec.height, spec.mpp, ctx=ctx)
        self.spec = spec

    def _load_image(self):
        return render_device(self.spec, 0, 0, self.w, self.h, ctx=self._ctx)


class ArrayWSI(DeviceWSI):
    """A host RGB array (or an image file PIL can open: the reference's ImageWSI case, core/wsi/image_wsi.py, mpp mandatory)
    uploaded to HBM once; afterwards every step of the path runs on the device copy."""

    def __init__(self, source, *, mpp: float, path: str | None = None, ctx: Context | None = None):
        if mpp is None or mpp <= 0:
            raise ValueError("mpp parameter is required for standard images")
        if isinstance(source, np.ndarray):
            arr = source
        else:
            from PIL import Image

            path = path or str(source)
            arr = np.asarray(Image.open(source).convert("RGB"))
        if arr.ndim != 3 or arr.shape[2] != 3 or arr.dtype != np.uint8:
            raise ValueError(f"expected an (H, W, 3) uint8 RGB array, got {arr.dtype} {arr.shape}")
        super().__init__(path or f"array_{arr.shape[1]}x{arr.shape[0]}", arr.shape[1], arr.shape[0], mpp, ctx=ctx)
        self._host = np.ascontiguousarray(arr)

    def _load_image(self):
        import torch

        pitch = (self.w * 3 + 15) // 16 * 16
        buf = torch.zeros((self.h, pitch), dtype=torch.uint8, device="cuda")
        buf[:, :self.w * 3] = torch.from_numpy(self._host.reshape(self.h, self.w * 3)).cuda()
        return buf, pitch


class DeviceWSILoader:
    """WSILoader (services/interfaces.py:34-40) for the runner: `.synth` descriptors -> SyntheticWSI, image files -> ArrayWSI."""

    def open(self, slide):
        from pathlib import Path

        p = Path(slide.path)
        if p.suffix.lower() == ".synth":
            from atlaspatch_b200.ref_backend import read_synth_descriptor
            from atlaspatch_b200.synthetic import make_spec

            d = read_synth_descriptor(p)
            mpp = slide.mpp if getattr(slide, "mpp", None) is not None else d["mpp"]
            return SyntheticWSI(make_spec(d["width"], d["height"], d["seed"], mpp=mpp), path=str(p))
        if getattr(slide, "mpp", None) is None:
            raise ValueError(f"{p.name}: mpp is required for standard images")
        return ArrayWSI(p, mpp=slide.mpp)
