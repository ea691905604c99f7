"""CPU: the per-GPU slide scheduler (atlaspatch_b200/runner.py) with the reference runner's semantics, on host doubles of the
device services (no GPU here): skip-existing, reuse of coordinates with missing feature sets, O_EXCL lock files, the failure
list (a slide that fails does not stop the run), --save-images, and a 2-rank gloo run that shards five slides over two ranks.

reference: orchestration/runner.py:106-181,202-306, services/feature_embedding.py:98-126,179-316, utils/features.py:37-71."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from atlaspatch_b200 import storage
from atlaspatch_b200.runner import (B200Runner, RunConfig, acquire_lock, existing_features, missing_features, patch_h5_path,
                                    patch_lock_path, release_lock)
from atlaspatch_b200.services import ExtractionConfig, ExtractionResult, Slide


class HostWSI:
    """Tiny host slide: deterministic pixels, single level, 20x."""

    def __init__(self, path, w=1024, h=768):
        self.path, self.w, self.h, self.mag, self.mpp, self.ds = str(path), w, h, 20, 0.5, [1.0]
        self.closed = False

    def get_size(self, lv=0):
        return self.w, self.h

    def metadata_attrs(self):
        return {"mpp": self.mpp, "magnification": self.mag}

    def extract(self, xy, lv, wh, mode="array"):
        rng = np.random.default_rng(abs(xy[0]) * 31 + abs(xy[1]))
        return rng.integers(0, 256, (wh[1], wh[0], 3), dtype=np.uint8)

    def cleanup(self):
        self.closed = True


class Loader:
    def __init__(self, fail=()):
        self.fail, self.opened = set(fail), []

    def open(self, slide):
        self.opened.append(slide.stem)
        if slide.stem in self.fail:
            raise OSError(f"cannot open {slide.path.name}")
        return HostWSI(slide.path)


class Seg:
    def __init__(self, fail=()):
        self.fail = set(fail)

    def segment_thumbnail(self, wsi):
        if Path(wsi.path).stem in self.fail:
            raise RuntimeError("segmentation exploded")
        m = np.zeros((48, 64), np.float32)
        m[8:40, 8:56] = 1.0
        return type("Mask", (), {"data": m})()


class Extraction:
    """extract(wsi, mask, slide=) through the CPU oracle of the coordinate path (test double of B200PatchExtractionService)."""

    def __init__(self, cfg):
        self.cfg = cfg

    def extract(self, wsi, mask, *, slide):
        from oracle import coords as oc

        c = oc.coords_from_mask(mask, level0_wh=wsi.get_size(), src_mag=wsi.mag, target_mag=self.cfg.target_magnification,
                                patch_size=self.cfg.patch_size, step_size=self.cfg.step_size, tissue_thresh=self.cfg.tissue_threshold)
        return ExtractionResult(slide=slide, h5_path=None, num_patches=int(c.shape[0]), coords=c, patch_size_level0=self.cfg.patch_size)


class Enc:
    def __init__(self, name, dim, fail=False):
        self.name, self.embedding_dim, self.input_patch, self.fail, self.cleaned = name, dim, 64, fail, False

    def extract_batch(self, patches, *, batch_size=None):
        if self.fail:
            raise RuntimeError("encoder exploded")
        if not len(patches):
            return np.empty((0, self.embedding_dim), np.float32)
        return np.stack([np.full(self.embedding_dim, float(p[0, 0, 0]), np.float32) for p in patches])

    def cleanup(self):
        self.cleaned = True


def _cfg(out, extractors=("enc_a",), **kw):
    return RunConfig(output_root=out, extraction=ExtractionConfig(patch_size=64, target_magnification=20, step_size=64),
                     feature_extractors=list(extractors), **kw)


def _runner(cfg, loader=None, seg=None, builders=None):
    builders = builders if builders is not None else {"enc_a": lambda: Enc("enc_a", 8), "enc_b": lambda: Enc("enc_b", 4)}
    return B200Runner(cfg, segmentation=seg or Seg(), extraction=Extraction(cfg.extraction), wsi_loader=loader or Loader(),
                      extractor_builders=builders)


def _slides(tmp_path, names):
    out = []
    for n in names:
        p = tmp_path / "slides" / f"{n}.tif"
        p.parent.mkdir(exist_ok=True)
        p.write_bytes(b"x" * (100 + 10 * len(n)))
        out.append(Slide(p, mpp=0.5))
    return out


def _read(h5_path):
    h5 = storage._h5py()
    with h5.File(str(h5_path), "r") as f:
        return {"coords": f["coords"][...], "attrs": dict(f.attrs.items()),
                "features": {k: v[...] for k, v in (f["features"].items() if "features" in f else [])}}


def test_full_run_then_skip_then_reuse_for_a_missing_feature_set(tmp_path):
    slides = _slides(tmp_path, ["a", "b", "c"])
    cfg = _cfg(tmp_path / "out")
    loader = Loader()
    results, failures = _runner(cfg, loader).run(slides)
    assert failures == [] and [r.slide.stem for r in results] == ["a", "b", "c"] and all(r.num_patches > 0 for r in results)
    d = _read(patch_h5_path(slides[0], cfg))
    assert d["features"]["enc_a"].shape == (results[0].num_patches, 8) and int(d["attrs"]["num_patches"]) == results[0].num_patches
    assert d["attrs"]["filename"] == "a.tif" and float(d["attrs"]["mpp"]) == 0.5
    assert not list((tmp_path / "out" / "patches").glob("*.lock"))
    # second run: everything complete -> nothing is opened, nothing is returned (runner.py:134-139)
    loader2 = Loader()
    results2, failures2 = _runner(cfg, loader2).run(slides)
    assert results2 == [] and failures2 == [] and loader2.opened == []
    # third run asks for one more encoder: coordinates are reused (same rows, no re-segmentation), only enc_b is embedded
    cfg3 = _cfg(tmp_path / "out", extractors=("enc_a", "enc_b"))
    before = _read(patch_h5_path(slides[1], cfg3))
    results3, failures3 = _runner(cfg3, seg=Seg(fail={"a", "b", "c"})).run(slides)      # segmentation would fail if it were called
    assert failures3 == [] and len(results3) == 3 and results3[0].metadata["feature_sets"] == ["enc_a", "enc_b"]
    after = _read(patch_h5_path(slides[1], cfg3))
    assert np.array_equal(after["coords"], before["coords"]) and np.array_equal(after["features"]["enc_a"], before["features"]["enc_a"])
    assert after["features"]["enc_b"].shape == (after["coords"].shape[0], 4)
    # --force: skip_existing off -> everything is redone
    loader4 = Loader()
    results4, _ = _runner(_cfg(tmp_path / "out", skip_existing=False), loader4).run(slides[:1])
    assert len(results4) == 1 and loader4.opened == ["a"]


def test_failures_are_collected_and_the_run_continues(tmp_path):
    slides = _slides(tmp_path, ["ok1", "noopen", "noseg", "ok2"])
    cfg = _cfg(tmp_path / "out", extractors=("enc_a", "bad"))
    runner = _runner(cfg, Loader(fail={"noopen"}), Seg(fail={"noseg"}),
                     builders={"enc_a": lambda: Enc("enc_a", 8), "bad": lambda: (_ for _ in ()).throw(RuntimeError("no weights"))})
    results, failures = runner.run(slides)
    assert [r.slide.stem for r in results] == ["ok1", "ok2"]
    by = {}
    for s, e in failures:
        by.setdefault(s.stem, []).append(str(e))
    assert "cannot open" in by["noopen"][0] and "segmentation exploded" in by["noseg"][0]
    assert by["ok1"] == ["no weights"] and by["ok2"] == ["no weights"]          # one entry per (slide, encoder that could not be built)
    assert set(_read(patch_h5_path(slides[0], cfg))["features"]) == {"enc_a"}
    assert not patch_h5_path(slides[1], cfg).exists() and not patch_h5_path(slides[2], cfg).exists()
    assert not list((tmp_path / "out" / "patches").glob("*.lock"))               # every lock released, also on the failure paths
    # an encoder that fails while embedding: the failure is recorded, no partial dataset is left behind
    cfg2 = _cfg(tmp_path / "out2", extractors=("boom",))
    results2, failures2 = _runner(cfg2, builders={"boom": lambda: Enc("boom", 4, fail=True)}).run(slides[:1])
    assert len(results2) == 1 and "encoder exploded" in str(failures2[0][1])
    assert _read(patch_h5_path(slides[0], cfg2))["features"] == {}


def test_lock_files_and_partial_feature_sets(tmp_path):
    slides = _slides(tmp_path, ["locked", "free"])
    cfg = _cfg(tmp_path / "out")
    lock = patch_lock_path(slides[0], cfg)
    fd = acquire_lock(lock, slides[0])
    assert fd is not None and acquire_lock(lock, slides[0]) is None and b"pid=" in lock.read_bytes()
    results, failures = _runner(cfg).run(slides)                                  # "locked by another process": skipped, no failure
    assert [r.slide.stem for r in results] == ["free"] and failures == [] and lock.exists()
    release_lock(fd, lock)
    assert not lock.exists()
    # a feature dataset whose row count differs from num_patches does not count as present (utils/features.py:50-57)
    h5 = patch_h5_path(slides[1], cfg)
    n = results[0].num_patches
    hl = storage._h5py()
    with hl.File(str(h5), "a") as f:
        f["features"].create_dataset("partial", data=np.zeros((n - 1, 3), np.float32))
    assert existing_features(h5, n) == {"enc_a"} and existing_features(h5, None) == {"enc_a", "partial"}
    assert missing_features(h5, ["ENC_A", "partial"], n) == ["partial"]


def test_save_images_writes_one_png_per_row(tmp_path):
    from PIL import Image

    slides = _slides(tmp_path, ["img"])
    cfg = _cfg(tmp_path / "out", extractors=(), save_images=True)
    results, failures = _runner(cfg).run(slides)
    assert failures == [] and results[0].image_dir == tmp_path / "out" / "images" / "img"
    files = sorted(p.name for p in results[0].image_dir.iterdir())
    assert len(files) == results[0].num_patches
    x, y = results[0].coords[0, :2]
    assert f"img_x{x}_y{y}.png" in files
    assert np.array_equal(np.asarray(Image.open(results[0].image_dir / f"img_x{x}_y{y}.png")), HostWSI("img").extract((int(x), int(y)), 0, (64, 64)))


def _two_rank_worker(rank, world, port, root):
    import torch.distributed as dist

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = Path(root)
    slides = [Slide(root / "slides" / f"{n}.tif", mpp=0.5) for n in ["s0", "s1", "s2", "s3", "s4"]]
    cfg = _cfg(root / "out")
    loader = Loader(fail={"s3"})
    results, failures = _runner(cfg, loader).run(slides)
    assert sorted(r.slide.stem for r in results) == ["s0", "s2", "s4"], [r.slide.stem for r in results]     # s1 pre-existing, s3 fails
    assert [s.stem for s, _ in failures] == ["s3"] and "cannot open" in failures[0][1]
    (root / f"opened_{rank}.txt").write_text(",".join(loader.opened))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_five_slides(tmp_path):
    """gloo, world_size 2: s1 is complete from an earlier run (skipped), s3 cannot be opened (failure), the other three are
    processed by the rank the LPT assignment gives them; both ranks end up with the same gathered (results, failures)."""
    import torch.multiprocessing as mp

    slides = _slides(tmp_path, ["s0", "s1", "s2", "s3", "s4"])
    cfg = _cfg(tmp_path / "out")
    _runner(cfg).run([slides[1]])                                                 # the earlier run
    port = 29500 + os.getpid() % 2000
    mp.spawn(_two_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    opened = [set(filter(None, (tmp_path / f"opened_{r}.txt").read_text().split(","))) for r in range(2)]
    assert opened[0] | opened[1] == {"s0", "s2", "s3", "s4"} and not (opened[0] & opened[1]) and opened[0] and opened[1]
    for n in ("s0", "s2", "s4"):
        d = _read(tmp_path / "out" / "patches" / f"{n}.h5")
        assert d["features"]["enc_a"].shape[0] == d["coords"].shape[0] > 0
