#!/bin/bash
# TEST INFRASTRUCTURE: installs the UNMODIFIED reference (AtlasPatch, pure Python) into baseline/_ref so that it travels to the
# GPU box (git-ignored, not gpurun-ignored).  Users: bench.py --impl reference (cpu_baseline.kind "reference"), and the
# reference-in-the-loop tests (tests/test_ref_in_loop.py, tests/test_gpu_ref_in_loop.py).  No reference source enters the git history.
#   --no-deps: the reference's dependencies that exist in this image are used as they are; the missing ones (openslide, h5py,
#   hydra, omegaconf, sam2, matplotlib, timm) are stubbed by oracle/refimport.py and never touched on the path.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/baseline/_ref"
[ -d "$REF/atlas_patch" ] || { echo "make_ref: no reference at $REF" >&2; exit 0; }
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/src"            # the build writes egg-info into the source tree; /root/reference is read-only
rm -rf "$DST"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$TMP/src" \
  || { echo "make_ref: pip install failed, copying the package directory instead" >&2; mkdir -p "$DST"; cp -r "$REF/atlas_patch" "$DST/"; }
rm -rf "$TMP"
echo "make_ref: installed $(ls "$DST" | tr '\n' ' ')into $DST"
