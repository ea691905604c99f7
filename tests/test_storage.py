"""CPU: the H5 writer (a10 / a14).  The layout logic (names, dtypes, shapes, chunking, attributes, tmp-then-move, atomic
replace) is checked on real files: through h5py where it is installed, else through atlaspatch_b200.h5lite (the in-tree HDF5
subset writer / reader; neither h5py nor libhdf5 exists in the build image) -- and against the file the reference's OWN,
unmodified H5PatchWriter produces for the same rows through the same HDF5 layer.  An in-memory double is kept only to inject
failures."""
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
    return storage._h5py(), True


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
    a = dict(f.attrs.items())
    assert {k: int(a[k]) for k in ("patch_size", "patch_size_level0", "level0_magnification", "target_magnification", "overlap",
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


def test_same_file_as_the_reference_writer(tmp_path):
    """The reference's unmodified H5PatchWriter.write_coords + append_features (services/storage.py:106-161,250-337) and
    storage.write_coords + append_features, both through the same HDF5 layer: identical datasets, dtypes, chunking and attributes."""
    from oracle import refimport

    if not refimport.reference_available():
        pytest.skip("reference not available")
    refimport.import_reference()
    from atlas_patch.services.storage import H5PatchWriter

    h5 = storage._h5py()
    n, d = 700, 24
    rng = np.random.default_rng(1)
    coords = np.concatenate([rng.integers(0, 50000, (n, 2)), np.full((n, 2), 512), np.zeros((n, 1))], 1).astype(np.int32)
    feats = rng.standard_normal((n, d)).astype(np.float32)
    extra = {"filename": "slideB.svs", "mpp": 0.25, "magnification": 40, "vendor": "aperio", "props": {"a": 1}}
    ref_path, our_path = tmp_path / "ref.h5", tmp_path / "our.h5"
    w = H5PatchWriter(chunk_rows=256, patch_size=256, patch_size_level0=512, level0_mag=40, target_mag=20, level0_wh=(50000, 40000),
                      overlap=64, slide_stem="slideB", wsi_path="/d/slideB.svs", extra_file_attrs=extra)
    total, _ = w.write_coords(ref_path, ((int(x), int(y), int(rw), int(rh), int(lv), None) for x, y, rw, rh, lv in coords), batch=256)
    assert total == n
    patches = [np.zeros((1, 1, 3), np.uint8)] * n
    it = iter(range(0, n, 32))
    w.append_features(output_path=ref_path, entries=((0, 0, 0, 0, 0, p) for p in patches), feature_name="enc",
                      feature_fn=lambda buf: feats[(s := next(it)):s + len(buf)], feature_attrs={"name": "enc", "embedding_dim": d},
                      feature_batch=32, expected_total=n)
    storage.write_coords(our_path, coords, slide_stem="slideB", wsi_path="/d/slideB.svs", patch_size=256, patch_size_level0=512,
                         level0_mag=40, target_mag=20, level0_wh=(50000, 40000), step_size=192, write_batch=256, extra_file_attrs=extra,
                         h5=h5)
    storage.append_features(our_path, "enc", feats, feature_batch=32, expected_total=n, h5=h5)
    with h5.File(str(ref_path), "r") as fr, h5.File(str(our_path), "r") as fo:
        assert sorted(fr.keys()) == sorted(fo.keys()) == ["coords", "features", "passports"]
        for key in ("coords", "passports", "features/enc"):
            a, b = fr[key], fo[key]
            assert a.shape == b.shape and a.dtype == b.dtype and a.chunks == b.chunks and a.maxshape == b.maxshape, key
            assert np.array_equal(a[...], b[...]), key
        ar, ao = dict(fr.attrs.items()), dict(fo.attrs.items())
        assert sorted(ar) == sorted(ao)
        for k in ar:
            if k != "creation_date":
                assert type(ar[k]) is type(ao[k]) and ar[k] == ao[k], k
        assert ao["props"] == '{"a": 1}' and ao["overlap"] == 64
