"""CPU: the UNMODIFIED reference's services run end to end in this harness (oracle/ref_pipeline.py) -- extraction -> H5 -> embedding
with the reference's own CPU extractor -> H5 -- and reproduce the committed golden vectors.  This is the arm bench.py --impl
reference times, and the control for tests/test_gpu_ref_in_loop.py, which swaps the B200 backend / plug-in into the same calls."""
import numpy as np
import pytest

from tests.cases import COORD_CASES, FEATURE_CASE, build_mask, case_spec


@pytest.fixture(scope="module")
def rp():
    from oracle import ref_pipeline

    if not ref_pipeline.reference_available():
        pytest.skip("reference not available (neither /root/reference nor baseline/_ref)")
    ref_pipeline.import_reference()
    return ref_pipeline


def test_reference_extract_writes_the_golden_coords(rp, tmp_path, golden_dir):
    case = COORD_CASES[0]                                    # BASELINE.json configs[0]: 8192^2, 256 px, stride 256
    spec = case_spec(case)
    wsi = rp.host_synthetic_wsi_class()(spec)
    extraction, _ = rp.reference_services(tmp_path, patch_size=case["patch"], target_mag=case["target_mag"], step_size=case["step"],
                                          tissue_threshold=case["tissue_thresh"])
    res = extraction.extract(wsi, build_mask(case, spec), slide=rp.reference_slide(wsi.path, mpp=spec.mpp))
    gold = np.load(golden_dir / f"coords_{case['name']}.npz")["coords"]
    got = rp.read_h5(res.h5_path)
    assert res.num_patches == gold.shape[0] and np.array_equal(got["coords"], gold)
    assert got["passports"][0].decode().startswith(f"{res.slide.stem}__x{gold[0, 0]}_y{gold[0, 1]}_rw256_rh256_lv0_mag20_tmag20_total")
    a = got["attrs"]
    assert int(a["num_patches"]) == gold.shape[0] and int(a["patch_size"]) == 256 and float(a["mpp"]) == 0.5
    assert a["filename"] == res.slide.path.name and int(a["magnification"]) == 20


def test_reference_embedding_loop_reproduces_golden_features(rp, tmp_path, golden_dir):
    from atlaspatch_b200.synthetic import make_spec
    from oracle.weights import vit_state_dict

    g = np.load(golden_dir / "vit_b_16_feats.npz")
    s = FEATURE_CASE["slide"]
    spec = make_spec(s["width"], s["height"], s["seed"], mpp=s["mpp"])
    wsi = rp.host_synthetic_wsi_class()(spec)
    rng = np.random.default_rng(77)                           # the coordinates of tests.cases.feature_patches()
    xy = [(int(rng.integers(0, spec.width - 256)), int(rng.integers(0, spec.height - 256))) for _ in range(FEATURE_CASE["n"])]
    coords = np.asarray([[x, y, 256, 256, 0] for x, y in xy], dtype=np.int32)[:6]
    sd = vit_state_dict("vit_b_16", seed=FEATURE_CASE["weight_seed"])
    extraction, embedding = rp.reference_services(tmp_path, patch_size=256, target_mag=20,
                                                  extractors={"vit_b_16": rp.reference_vit_builder(sd, num_workers=0)}, feature_batch=4)
    slide = rp.reference_slide(wsi.path, mpp=spec.mpp)
    res = rp.write_reference_coords(embedding, wsi, slide, coords, patch_size_level0=256)
    extractor = embedding.registry.create("vit_b_16")
    embedding._embed_with_extractor(result=res, wsi=wsi, extractor=extractor)
    got = rp.read_h5(res.h5_path)
    assert list(got["features"]) == ["vit_b_16"] and got["features"]["vit_b_16"].shape == (6, 768)
    ref = g["feats"][:6]
    rel = np.linalg.norm(got["features"]["vit_b_16"] - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 2e-5, rel                               # same code, same weights, same patches: float noise only
    assert res.metadata["feature_sets"] == ["vit_b_16"]
    assert not list(res.h5_path.parent.glob("*.lock"))         # the feature lock was released
