"""Seam #3: the HBM-resident slide as a backend of the reference's own WSIFactory.

    from atlaspatch_b200.ref_backend import register_synthetic_backend
    register_synthetic_backend()                       # WSIFactory.register("synthetic", ...) + map_extension(".synth", ...)
    wsi = WSIFactory.load("slide.synth")               # reference: core/wsi/wsi_factory.py:41-94

The class is a real subclass of the reference's `IWSI` ABC (core/wsi/iwsi.py:9-124): `_setup / _extract_mpp / _extract_mag /
extract / get_size / get_thumb / cleanup` are implemented on top of `slide.SyntheticWSI`, `get_thumbnail_at_power`
(iwsi.py:246-323) is overridden to run on the device (ap_thumbnail_resize: same pixels as the inherited cv2 path, without the
whole-level host read), and `device_image` / `pitch` hand the level-0 pointer to the B200 services.  The reference is imported
lazily: the package itself never needs `atlas_patch`.

A `.synth` file is a small JSON descriptor {"width", "height", "seed", "mpp"} of a synthetic slide (synthetic.py).
"""
from __future__ import annotations

import json
import os
from pathlib import Path

from atlaspatch_b200.synthetic import make_spec

BACKEND_NAME = "synthetic"
EXTENSION = ".synth"


def write_synth_descriptor(path, width: int, height: int, seed: int = 0, mpp: float = 0.5) -> Path:
    path = Path(path)
    path.write_text(json.dumps({"width": int(width), "height": int(height), "seed": int(seed), "mpp": float(mpp)}))
    return path


def read_synth_descriptor(path) -> dict:
    d = json.loads(Path(path).read_text())
    for k in ("width", "height"):
        if int(d.get(k, 0)) <= 0:
            raise ValueError(f"{path}: '{k}' must be a positive integer")
    return {"width": int(d["width"]), "height": int(d["height"]), "seed": int(d.get("seed", 0)), "mpp": float(d.get("mpp", 0.5))}


def make_backend_class(iwsi_base):
    """Subclass of the given IWSI ABC (the reference's, or a compatible one) around slide.SyntheticWSI."""
    from atlaspatch_b200.slide import SyntheticWSI

    class SyntheticBackend(iwsi_base):
        """Single-level synthetic slide living in HBM, behind the reference's IWSI contract."""

        def __init__(self, path, mpp=None, **kwargs):
            super().__init__(path=path, mpp=mpp)
            self._impl = None

        # ---- IWSI abstract methods ----
        def _setup(self) -> None:
            d = read_synth_descriptor(self.path)
            mpp = float(self._mpp_manual) if self._mpp_manual is not None else d["mpp"]
            self._impl = SyntheticWSI(make_spec(d["width"], d["height"], d["seed"], mpp=mpp), path=os.fspath(self.path))
            self.w, self.h = self._impl.w, self._impl.h
            self.nlvl, self.ds, self.dims = 1, [1.0], [(self.w, self.h)]
            self.meta = {"format": "synthetic", "seed": d["seed"]}
            self.mpp = self._extract_mpp()
            self.mag = self._extract_mag()

        def _extract_mpp(self):
            return self.validate_mpp(float(self._impl.spec.mpp), source="synthetic descriptor")

        def _extract_mag(self):
            return self._infer_mag(self.mpp) if self.mpp is not None else None

        def extract(self, xy, lv, wh, *, mode="array"):
            self._ensure_loaded()
            if int(lv) != 0:
                raise ValueError(f"synthetic slides have a single level (asked for level {lv})")
            return self._impl.extract(xy, 0, wh, mode=mode)

        def get_size(self, lv: int = 0):
            self._ensure_loaded()
            return self.w, self.h

        def get_thumb(self, max_hw):
            self._ensure_loaded()
            return self._impl.get_thumb(max_hw)

        def cleanup(self) -> None:
            if self._impl is not None:
                self._impl.cleanup()

        # ---- device fast paths ----
        def get_thumbnail_at_power(self, *, power: float = 1.25, interpolation: str = "optimise"):
            self._ensure_loaded()
            if interpolation not in ("optimise", "area"):
                return super().get_thumbnail_at_power(power=power, interpolation=interpolation)
            return self._impl.get_thumbnail_at_power(power=power, interpolation=interpolation)

        @property
        def device_image(self):
            self._ensure_loaded()
            return self._impl.device_image

        @property
        def pitch(self) -> int:
            self._ensure_loaded()
            return self._impl.pitch

        @property
        def spec(self):
            self._ensure_loaded()
            return self._impl.spec

    SyntheticBackend.__name__ = "SyntheticBackend"
    return SyntheticBackend


def register_synthetic_backend(wsi_factory=None, iwsi_base=None):
    """Registers the backend with the reference's WSIFactory (imported from `atlas_patch` unless given) and returns the class."""
    if wsi_factory is None:
        from atlas_patch.core.wsi.wsi_factory import WSIFactory as wsi_factory   # noqa: N813
    if iwsi_base is None:
        from atlas_patch.core.wsi.iwsi import IWSI as iwsi_base                   # noqa: N813
    cls = make_backend_class(iwsi_base)
    wsi_factory.register(BACKEND_NAME, cls)
    wsi_factory.map_extension(EXTENSION, BACKEND_NAME)
    return cls
