"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the UNMODIFIED reference (`/root/reference/atlas_patch`) in this build
container so that its own functions can (a) pin the oracle restatements in
`oracle/` and (b) generate the golden vectors committed under `tests/golden/`.

The reference needs packages this image lacks (openslide, h5py, hydra,
omegaconf, sam2, matplotlib, timm); none of them is touched by the functions we
call on the hot path (coordinate extraction, thumbnailing, the torchvision ViT
extractor), so they are satisfied with empty stub modules.  On the GPU box
`/root/reference` does not exist; there the same unmodified package is found under
`baseline/_ref` when oracle/make_ref.sh has installed it (bench.py --impl reference and the
reference-in-the-loop GPU tests use it; they skip / fall back to the port when it is absent).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# baseline/_ref: the unmodified reference installed by oracle/make_ref.sh (pip install --target; git-ignored, travels to the GPU box)
INSTALLED_REF = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _pick_root() -> str:
    env = os.environ.get("ATLAS_REF")
    if env:
        return env
    if os.path.isdir("/root/reference/atlas_patch"):
        return "/root/reference"
    return INSTALLED_REF


REFERENCE_ROOT = _pick_root()

_STUBS = {
    "openslide": {"OpenSlide": type("OpenSlide", (), {}), "OpenSlideError": Exception,
                  "OpenSlideUnsupportedFormatError": Exception, "PROPERTY_NAME_MPP_X": "openslide.mpp-x",
                  "PROPERTY_NAME_MPP_Y": "openslide.mpp-y", "PROPERTY_NAME_OBJECTIVE_POWER": "openslide.objective-power"},
    "h5py": {"File": None, "Dataset": type("Dataset", (), {}), "Group": type("Group", (), {}),
             "string_dtype": lambda *a, **k: None},
    "hydra": {},
    "hydra.utils": {"instantiate": None},
    "omegaconf": {"OmegaConf": None},
    "sam2": {},
    "sam2.sam2_image_predictor": {"SAM2ImagePredictor": type("SAM2ImagePredictor", (), {})},
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "matplotlib.cm": {},
    "timm": {"create_model": None},
    "timm.layers": {"to_2tuple": lambda x: (x, x)},
    "timm.data": {"resolve_data_config": None},
    "timm.data.transforms_factory": {"create_transform": None},
}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "atlas_patch"))


def install_stubs() -> None:
    if "h5py" not in sys.modules:
        try:
            importlib.import_module("h5py")
        except Exception:
            # no libhdf5 in this image: the reference's h5py calls (utils/h5.py, services/storage.py, utils/features.py,
            # orchestration/runner.py) run on the in-tree HDF5 subset implementation, which exposes the same API
            from atlaspatch_b200 import h5lite

            sys.modules["h5py"] = h5lite
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        mod.__path__ = []  # behave like a package so sub-imports resolve
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            if parent in sys.modules:
                setattr(sys.modules[parent], child, mod)


def import_reference():
    """Return the imported `atlas_patch` package of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("atlas_patch")
