"""CPU tests of the synthetic slide definition (host generator)."""
import numpy as np

from atlaspatch_b200.synthetic import make_spec, render_region_host, truth_mask


def test_regions_are_consistent_and_padded():
    spec = make_spec(2048, 1024, seed=9)
    full = render_region_host(spec, 0, 0, 2048, 1024)
    sub = render_region_host(spec, 300, 200, 257, 129)
    assert np.array_equal(sub, full[200:329, 300:557])
    over = render_region_host(spec, -10, 1000, 64, 64)
    assert (over[:, :10] == 0).all() and (over[24:] == 0).all()
    assert np.array_equal(over[:24, 10:], full[1000:1024, 0:54])


def test_seeded_and_nontrivial():
    a, b = make_spec(4096, 4096, seed=1), make_spec(4096, 4096, seed=1)
    assert a == b and a != make_spec(4096, 4096, seed=2)
    m = truth_mask(a)
    assert m.shape == (256, 256) and 0.05 < m.mean() < 0.9
    assert len(a.holes) >= 1 and 3 <= len(a.blobs) <= 6
    # tissue pixels are darker / more saturated than background
    px = render_region_host(a, 0, 0, 4096, 16)[0]
    assert px.min() >= 60 and px.max() <= 239
