"""CPU: the H5 writer (a10 / a14) against an in-memory double of the h5py calls it makes.  Neither h5py nor libhdf5 exists in the
build image, so what is checked here is the layout logic (names, dtypes, shapes, chunking, attributes, tmp-then-move, atomic
replace) -- not the bytes on disk; with h5py installed the same test runs against the real library."""
import os
from pathlib import Path

import numpy as np
import pytest

from atlaspatch_b200 import storage


class _Dataset:
    def __init__(self, shape, maxshape, chunks, dtype):
        self.maxshape, self.chunks, self.dtype = maxshape, chunks, np.dtype(dtype)
        self.data = np.zeros(shape, dtype=self.dtype)

    @property
    def shape(self):
        return self.data.shape

    def resize(self, size, axis=None):
        new = list(self.data.shape)
        if axis is None:
            new = list(size)
        else:
            new[axis] = size
        assert all(m is None or n <= m for n, m in zip(new, self.maxshape))
        grown = np.zeros(new, dtype=self.dtype)
        grown[tuple(slice(0, min(a, b)) for a, b in zip(self.data.shape, new))] = self.data[tuple(slice(0, min(a, b)) for a, b in zip(self.data.shape, new))]
        self.data = grown

    def __setitem__(self, key, value):
        self.data[key] = value


class _Group(dict):
    def __init__(self):
        super().__init__()
        self.attrs = {}

    def create_dataset(self, name, shape, maxshape, chunks, dtype):
        assert name not in self
        self[name] = _Dataset(shape, maxshape, chunks, dtype)
        return self[name]

    def require_group(self, name):
        return self.setdefault(name, _Group())

    def move(self, src, dst):
        assert dst not in self
        self[dst] = self.pop(src)


class _FakeH5:
    """h5.File(path, mode): contents live in a dict keyed by a token stored in the (otherwise empty) file at `path`, so that
    os.replace of the temporary file carries them along like a real file."""
    store: dict = {}

    class File(_Group):
        def __init__(self, path, mode):
            super().__init__()
            if mode in ("a", "r") and os.path.exists(path):
                old = _FakeH5.store[Path(path).read_text()]
                self.update(old)
                self.attrs = old.attrs
                _FakeH5.store[Path(path).read_text()] = self
            else:
                key = f"fake-{len(_FakeH5.store)}"
                Path(path).write_text(key)
                _FakeH5.store[key] = self

        def close(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False


def _real_or_fake():
    try:
        import h5py
    except ImportError:
        return _FakeH5, False
    if not (hasattr(h5py, "h5f") and callable(getattr(h5py, "File", None))):   # oracle/refimport.py's stub module, not the library
        return _FakeH5, False
    return h5py, True


def test_passports_follow_the_reference_format():
    c = np.array([[10, 20, 256, 256, 0], [266, 20, 512, 512, 0]], dtype=np.int32)
    p = storage.passports("slideA", c, 40, 20)
    assert p.dtype == np.dtype("S160") and p.shape == (2,)
    assert p[0] == b"slideA__x10_y20_rw256_rh256_lv0_mag40_tmag20_total2"
    assert storage.passports("s", c[:1], 0, 0)[0] == b"s__x10_y20_rw256_rh256_lv0_magna_tmagna_total1"


def test_layout_names_dtypes_chunks_attrs_and_feature_move(tmp_path):
    h5, real = _real_or_fake()
    n = 20000
    coords = np.stack([np.arange(n) * 256, np.arange(n) * 3, np.full(n, 256), np.full(n, 256), np.zeros(n)], 1).astype(np.int32)
    out = tmp_path / "patches" / "slideA.h5"
    out.parent.mkdir()
    got = storage.write_coords(out, coords, slide_stem="slideA", wsi_path="/data/slideA.svs", patch_size=256, patch_size_level0=512,
                               level0_mag=40, target_mag=20, level0_wh=(80000, 60000), step_size=128, write_batch=8192, h5=h5)
    assert got == n and out.exists() and [p.name for p in out.parent.iterdir()] == ["slideA.h5"]   # tmp file replaced, nothing left
    feats = np.random.default_rng(0).standard_normal((n, 768)).astype(np.float32)
    assert storage.append_features(out, "vit_b_16", feats, feature_batch=32, expected_total=n, h5=h5) == n
    with pytest.raises(ValueError, match="already exists"):
        storage.append_features(out, "vit_b_16", feats, h5=h5)
    with pytest.raises(ValueError, match="do not match expected coords"):
        storage.append_features(out, "other", feats[:-1], expected_total=n, h5=h5)
    with pytest.raises(ValueError, match="2D array"):
        storage.append_features(out, "other", feats[0], h5=h5)

    f = h5.File(str(out), "r")
    dc, dp, df = f["coords"], f["passports"], f["features"]["vit_b_16"]
    assert dc.shape == (n, 5) and dc.dtype == np.int32 and dc.maxshape == (None, 5) and dc.chunks == (8192, 5)
    assert dp.shape == (n,) and dp.dtype == np.dtype("S160") and dp.chunks == (8192,)
    assert df.shape == (n, 768) and df.dtype == np.float32 and df.maxshape == (None, 768) and df.chunks == (32, 768)
    assert list(f["features"].keys()) == ["vit_b_16"]                                              # no __tmp_ dataset left behind
    data = (lambda d: d[...]) if real else (lambda d: d.data)
    assert np.array_equal(data(dc), coords) and np.array_equal(data(df), feats)
    assert data(dp)[-1] == f"slideA__x{(n - 1) * 256}_y{(n - 1) * 3}_rw256_rh256_lv0_mag40_tmag20_total{n}".encode()
    a = dict(f.attrs)
    assert {k: a[k] for k in ("patch_size", "patch_size_level0", "level0_magnification", "target_magnification", "overlap",
                              "level0_width", "level0_height", "num_patches", "passport_version")} == \
        dict(patch_size=256, patch_size_level0=512, level0_magnification=40, target_magnification=20, overlap=128,
             level0_width=80000, level0_height=60000, num_patches=n, passport_version=2)
    assert a["wsi_path"] == "/data/slideA.svs" and a["filename"] == "slideA.svs" and a["passport_format"] == storage.PASSPORT_FORMAT
    assert "creation_date" in a


def test_failed_write_leaves_no_partial_file(tmp_path):
    class Boom(_FakeH5):
        class File(_FakeH5.File):
            def create_dataset(self, *a, **k):
                raise RuntimeError("disk full")

    with pytest.raises(RuntimeError, match="disk full"):
        storage.write_coords(tmp_path / "x.h5", np.zeros((3, 5), np.int32), slide_stem="x", wsi_path="x.svs", patch_size=256,
                             patch_size_level0=256, level0_mag=20, target_mag=20, level0_wh=(10, 10), h5=Boom)
    assert list(tmp_path.iterdir()) == []
