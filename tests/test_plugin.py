"""CPU: the --feature-plugin hook (reference: models/patch/custom.py:92-146) registers builders without touching the GPU, a
builder refuses to run without a CUDA device (no CPU fallback), and -- with the reference's OWN loader and registry -- the
plug-in file loads by path from a foreign working directory."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
NAMES = {"b200_vit_b_16", "b200_vit_l_16", "b200_vit_b_32", "b200_vit_l_32", "b200_dinov2_small", "b200_dinov2_base",
         "b200_dinov2_large", "b200_dinov2_giant", "b200_midnight", "b200_phikon_v1", "b200_phikon_v2", "b200_hibou_b", "b200_hibou_l",
         "b200_openmidnight", "b200_plip", "b200_quilt_b_32", "b200_quilt_b_16", "b200_h_optimus_0",
         "b200_h_optimus_1", "b200_pathorchestra", "b200_prov_gigapath",
         "b200_clip_vit_b_32", "b200_clip_vit_b_16", "b200_clip_vit_l_14",
         "b200_uni_v1", "b200_uni_v2", "b200_h0_mini", "b200_lunit_vit_small_patch16_dino"}


def _reference():
    from oracle import refimport

    if not refimport.reference_available():
        pytest.skip("reference not available (neither /root/reference nor baseline/_ref)")
    refimport.import_reference()


def test_hook_registers_with_the_reference_registry():
    """models/patch/registry.py:11-44 + custom.py:113-146, unmodified: names lower-cased, duplicates rejected."""
    _reference()
    import torch
    from atlas_patch.models.patch.custom import register_feature_extractors_from_module
    from atlas_patch.models.patch.registry import PatchFeatureExtractorRegistry

    reg = PatchFeatureExtractorRegistry()
    plugin = ROOT / "atlaspatch_b200" / "plugin.py"
    register_feature_extractors_from_module(plugin, reg, device=torch.device("cuda"), dtype=torch.float16, num_workers=0)
    assert set(reg.available()) == NAMES
    with pytest.raises(ValueError):   # a second load collides with the names already registered
        register_feature_extractors_from_module(plugin, reg, device=torch.device("cuda"), dtype=torch.float16, num_workers=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):   # builders run lazily, at registry.create
        reg2 = PatchFeatureExtractorRegistry()
        register_feature_extractors_from_module(plugin, reg2, device=torch.device("cpu"), dtype=torch.float32, num_workers=0)
        reg2.create("b200_vit_b_16")


def test_plugin_loads_by_path_from_a_foreign_cwd(tmp_path):
    """The reference imports the plug-in with spec_from_file_location: the package root is NOT on sys.path (no PYTHONPATH,
    cwd elsewhere).  plugin.py bootstraps it."""
    code = (
        "import importlib.util, sys\n"
        f"spec = importlib.util.spec_from_file_location('plugin', r'{ROOT / 'atlaspatch_b200' / 'plugin.py'}')\n"
        "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
        "names = []\n"
        "class R:\n"
        "    def register(self, n, b): names.append(n)\n"
        "m.register_feature_extractors(registry=R(), device='cuda', dtype=None, num_workers=0)\n"
        "print(sorted(names))\n")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    res = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert eval(res.stdout.strip().splitlines()[-1]) == sorted(NAMES)


def test_builders_refuse_cpu_devices():
    from atlaspatch_b200.plugin import _build, _build_dinov2

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _build("vit_b_16", "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _build_dinov2("dinov2_large", "cpu", 224)
    from atlaspatch_b200.plugin import _HUB, _TIMM, _build_hub, _build_timm

    for name in _HUB:                       # refused before any hub / timm / open_clip import is attempted
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _build_hub(name, "cpu", None)
    for name in _TIMM:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _build_timm(name, "cpu", None)


def test_dtype_policy_matches_the_reference():
    """services/feature_embedding.py:28-39."""
    import torch

    from atlaspatch_b200.plugin import resolve_feature_dtype

    cases = [("cpu", "float16"), ("cpu", "float32"), ("cpu", "bfloat16"), ("cuda", "float16"), ("cuda:1", "bfloat16"), ("cuda", "nonsense")]
    want = {("cpu", "float16"): torch.float32, ("cpu", "float32"): torch.float32, ("cpu", "bfloat16"): torch.bfloat16,
            ("cuda", "float16"): torch.float16, ("cuda:1", "bfloat16"): torch.bfloat16, ("cuda", "nonsense"): torch.float32}
    for dev, prec in cases:
        assert resolve_feature_dtype(torch.device(dev), prec) == want[(dev, prec)]
    try:
        _reference()
    except pytest.skip.Exception:
        return
    from atlas_patch.services.feature_embedding import resolve_feature_dtype as ref_fn

    for dev, prec in cases:
        assert resolve_feature_dtype(torch.device(dev), prec) == ref_fn(torch.device(dev), prec)
