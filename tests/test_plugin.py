"""CPU: the --feature-plugin hook (reference: models/patch/custom.py:92-146) registers builders without touching the GPU, and a
builder refuses to run without a CUDA device (no CPU fallback)."""
import pytest


class _Registry:  # the two methods of PatchFeatureExtractorRegistry the hook uses (models/patch/registry.py:11-44)
    def __init__(self):
        self.builders = {}

    def register(self, name, builder):
        if name in self.builders:
            raise ValueError(f"Feature extractor '{name}' is already registered.")
        self.builders[name] = builder


def test_hook_registers_all_b200_encoders():
    from atlaspatch_b200.plugin import register_feature_extractors

    reg = _Registry()
    register_feature_extractors(reg, "cpu", None, 0)
    assert set(reg.builders) == {"b200_vit_b_16", "b200_vit_l_16", "b200_dinov2_large", "b200_dinov2_giant"}
    with pytest.raises(ValueError):
        register_feature_extractors(reg, "cpu", None, 0)   # duplicate names are rejected like the reference's registry does


def test_builders_refuse_cpu_devices():
    from atlaspatch_b200.plugin import _build, _build_dinov2

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _build("vit_b_16", "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _build_dinov2("dinov2_large", "cpu", 224)
