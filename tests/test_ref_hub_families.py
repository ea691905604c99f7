"""CPU: the oracle of the hub encoder families (oracle/hub_families.py) pinned against the UNMODIFIED reference's own extractor classes.

The reference's classes load their weights from the hub (no network here), so the hub calls -- and only those -- are replaced: the
`from_pretrained` / `timm.create_model` entry points return the oracle's seeded tiny model (and, for the families whose processor comes
from the hub, the processor object the oracle builds).  Everything else is the reference's code: the class constructor, its preprocess
closure / torchvision Compose, PatchDataset -> DataLoader -> forward_fn -> numpy of models/patch/base.py:76-107.  What this pins: the
preprocess each class builds, the forward_fn (which output, which tokens are averaged, in which order they are concatenated) and the
float32 conversion.  What stays restated from the public hub files: the processor settings of phikon / hibou / plip / quilt."""
import contextlib
from unittest import mock

import numpy as np
import pytest
import torch

from oracle import hub_families as hf


@pytest.fixture(scope="module")
def ref():
    from oracle import refimport

    if not refimport.reference_available():
        pytest.skip("reference not available (neither /root/reference nor baseline/_ref)")
    return refimport.import_reference()


def _patches(n=3, P=256):
    rng = np.random.default_rng(5)
    out = []
    for i in range(n):
        base = rng.integers(0, 256, (P // 16 + 1, P // 16 + 1, 3), dtype=np.uint8)
        img = np.kron(base, np.ones((16, 16, 1), dtype=np.uint8))[:P, :P]
        out.append((img.astype(np.int32) + rng.integers(-25, 25, img.shape)).clip(0, 255).astype(np.uint8))
    return out


class _Processor:
    """Stands in for the hub's processor object: the reference calls processor(images=pil, return_tensors="pt")["pixel_values"]."""

    def __init__(self, name):
        self.pre = hf.make_preprocess(name)

    def __call__(self, images=None, return_tensors="pt"):
        return {"pixel_values": self.pre(images).unsqueeze(0)}


@contextlib.contextmanager
def _hub(model, processor=None):
    """Replace the hub entry points the reference's classes call (phikon.py:41-46,90-93, midnight.py:44, hibou.py:51-54, plip.py:34-35,
    quilt.py:56-57) for the duration of one constructor."""
    import transformers

    ret_model = lambda *a, **k: model          # noqa: E731
    ret_proc = lambda *a, **k: processor       # noqa: E731
    with contextlib.ExitStack() as st:
        for cls in ("AutoModel", "ViTModel", "CLIPModel"):
            st.enter_context(mock.patch.object(getattr(transformers, cls), "from_pretrained", ret_model))
        for cls in ("AutoImageProcessor", "CLIPProcessor", "AutoProcessor"):
            st.enter_context(mock.patch.object(getattr(transformers, cls), "from_pretrained", ret_proc))
        yield


def _check(extractor, name, sd, want_dim):
    patches = _patches()
    got = extractor.extract_batch(patches, batch_size=2)                    # the reference's own loop, CPU fp32
    want = hf.extract_features(patches, sd, name)
    assert got.dtype == np.float32 and got.shape == want.shape == (3, want_dim)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), np.abs(got - want).max()


CPU = dict(device=torch.device("cpu"), dtype=torch.float32, num_workers=0)


def test_midnight_class(ref):
    from atlas_patch.models.patch.midnight import Midnight

    name = "midnight_test_tiny"
    sd = hf.state_dict(name, seed=11)
    with _hub(hf.build_model(name, sd)):
        ext = Midnight(**CPU)
    _check(ext, name, sd, 768)                                             # [class || mean of patch tokens] of the 384-wide tiny model


def test_phikon_classes(ref):
    from atlas_patch.models.patch.phikon import Phikon, PhikonV2

    for cls, name, dim in ((Phikon, "phikon_v1_test_tiny", 256), (PhikonV2, "phikon_v2_test_tiny", 256)):
        sd = hf.state_dict(name, seed=12)
        with _hub(hf.build_model(name, sd), _Processor(name)):
            ext = cls(**CPU)
        _check(ext, name, sd, dim)


def test_hibou_class(ref):
    from atlas_patch.models.patch.hibou import HibouEncoder

    name = "hibou_test_tiny"
    sd = hf.state_dict(name, seed=13)
    with _hub(hf.build_model(name, sd), _Processor(name)):
        ext = HibouEncoder(name="hibou_b", model_id="histai/hibou-B", embedding_dim=256, **CPU)
    _check(ext, name, sd, 256)                                             # pooler_output of the register-token model


class _Clip4x(torch.nn.Module):
    """transformers 4.x semantics of CLIPModel.get_image_features (what the reference's forward_fn expects, plip.py:56): the projected
    features as a tensor.  transformers >= 5 (this image) returns the vision ModelOutput with them in pooler_output."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def get_image_features(self, pixel_values=None):
        out = self.model.get_image_features(pixel_values=pixel_values)
        return out if isinstance(out, torch.Tensor) else out.pooler_output


def test_plip_and_quilt_classes(ref):
    from atlas_patch.models.patch.plip import PLIPExtractor
    from atlas_patch.models.patch.quilt import QuiltNet

    name = "plip_test_tiny"
    sd = hf.state_dict(name, seed=14)
    with _hub(_Clip4x(hf.build_model(name, sd)), _Processor(name)):
        ext = PLIPExtractor(**CPU)
    _check(ext, name, sd, 128)
    name = "quilt_b_16_test_tiny"
    sd = hf.state_dict(name, seed=15)
    with _hub(_Clip4x(hf.build_model(name, sd)), _Processor(name)):
        ext = QuiltNet(name="quilt_b_16", model_id="wisdomik/QuiltNet-B-16", **CPU)
    _check(ext, name, sd, 128)


def test_explicit_transforms_of_the_timm_and_hub_classes(ref):
    """hoptimus.py:14-31, gigapath.py:17-26, openmidnight.py:17-30, midnight.py:15-25 spell their torchvision transforms out: the oracle's
    preprocess objects must give the same tensors, and the integer restatements (what the CUDA kernels implement) the same pixels."""
    from PIL import Image

    from atlas_patch.models.patch import gigapath, hoptimus, midnight, openmidnight

    pairs = (("h_optimus_test_tiny", hoptimus._build_hoptimus_transform()), ("prov_gigapath_test_tiny", gigapath._build_preprocess()),
             ("openmidnight_test_tiny", openmidnight._build_preprocess()), ("midnight_test_tiny", midnight._build_preprocess()))
    for name, ref_tf in pairs:
        mine = hf.make_preprocess(name)
        for P in (224, 256, 512):
            pil = Image.fromarray(_patches(1, P)[0])
            want = ref_tf(pil)
            assert torch.equal(mine(pil), want), (name, P)
            # un-normalise the reference's tensor: it must be exactly the uint8 pixels of the integer restatement
            mean = torch.tensor(ref_tf.transforms[-1].mean).view(3, 1, 1)
            std = torch.tensor(ref_tf.transforms[-1].std).view(3, 1, 1)
            pix = torch.round((want * std + mean) * 255.0).permute(1, 2, 0).numpy().astype(np.uint8)
            assert np.array_equal(pix, hf.pixels(name, np.asarray(pil))), (name, P)


def test_pathorchestra_class_transform(ref):
    """pathorchestra.py:38-58 builds its transform inside the constructor; timm.create_model is the hub call replaced."""
    import timm
    from PIL import Image

    from atlas_patch.models.patch.pathorchestra import PathOrchestraEncoder

    name = "pathorchestra_test_tiny"
    sd = hf.state_dict(name, seed=16)
    model = hf.build_model(name, sd)

    class _Token(torch.nn.Module):                 # timm's num_classes = 0, global_pool "token": forward -> class token after the norm
        def forward(self, x):
            return model(pixel_values=x).last_hidden_state[:, 0]

    with mock.patch.object(timm, "create_model", lambda *a, **k: _Token(), create=True):
        ext = PathOrchestraEncoder(**CPU)
    pil = Image.fromarray(_patches(1, 300)[0])
    assert torch.equal(ext.preprocess(pil), hf.make_preprocess(name)(pil))
    _check(ext, name, sd, 256)


@pytest.mark.parametrize("module,cls,name,dim", [("hoptimus", "HOptimus0", "h_optimus_test_tiny", 384), ("hoptimus", "HOptimus1", "h_optimus_test_tiny", 384),
                                                 ("gigapath", "ProvGigaPathExtractor", "prov_gigapath_test_tiny", 384)])
def test_timm_loaded_classes(ref, module, cls, name, dim):
    """hoptimus.py:34-132, gigapath.py:29-66: timm.create_model (the hub call) returns the oracle's model behind timm's forward contract
    (num_classes 0, global_pool "token": the class token after the final norm); constructor, transform and loop are the reference's."""
    import importlib

    import timm

    sd = hf.state_dict(name, seed=17)
    model = hf.build_model(name, sd)

    class _Token(torch.nn.Module):
        def forward(self, x):
            return model(pixel_values=x).last_hidden_state[:, 0]

    klass = getattr(importlib.import_module(f"atlas_patch.models.patch.{module}"), cls)
    with mock.patch.object(timm, "create_model", lambda *a, **k: _Token(), create=True):
        ext = klass(**CPU)
    _check(ext, name, sd, dim)
