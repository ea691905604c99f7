"""CPU: bench.py's copies of the golden cases stay in step with tests/cases.py, and its argument defaults follow the contract."""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_golden_cases_match_tests_cases():
    from tests.cases import DINOV2_CASES, dinov2_coords

    b = _bench()
    assert b.GOLDEN_CASES == DINOV2_CASES
    for name, case in b.GOLDEN_CASES.items():       # the row generator inside bench._golden_check
        rng = np.random.default_rng(99)
        P, sl = case["patch"], case["slide"]
        xy = np.stack([rng.integers(0, sl["width"] - P, case["n"]), rng.integers(0, sl["height"] - P, case["n"])], 1)
        xy[-1] = (sl["width"] - P // 2, sl["height"] - P // 3)
        rows = np.concatenate([xy, np.full((case["n"], 2), P), np.zeros((case["n"], 1))], 1).astype(np.int32)
        assert np.array_equal(rows, dinov2_coords(name))


def test_defaults_and_no_oracle_import_on_the_gpu_arm():
    b = _bench()
    old = sys.argv
    sys.argv = ["bench.py"]
    try:
        a = b.parse_args()
    finally:
        sys.argv = old
    assert a.gpus == 1 and a.warmup >= 3 and a.e2e_steps >= 10 and a.aux == "c2,c3,c4"
    src = (ROOT / "bench.py").read_text()
    gpu_arm = src[src.index("def host_patches_from_slide"):]
    gpu_arm = gpu_arm[:gpu_arm.index("    cpu_baseline = None")]
    assert "oracle" not in gpu_arm.replace("oracle/", "")    # the product arm never imports the oracle (comments may name its files)
