"""CPU: the tap tables the CUDA resizing preprocesses run from (host-built in csrc/preprocess_resize.cu, exported as ap_resize_tap_tables)
against the coefficient arithmetic of the libraries they restate -- oracle/resize_aa.py, which other CPU tests pin bit for bit against
torch (ATen uint8 antialias bicubic / bilinear) and Pillow (BILINEAR / BICUBIC) themselves.  No device is needed for this entry."""
import ctypes as C

import numpy as np
import pytest

from oracle import resize_aa as ra


@pytest.fixture(scope="module")
def lib():
    from atlaspatch_b200._lib import load_library

    return load_library()


def _tables(lib, filt, n_in, n_out, image, cap=64):
    tmin, tcnt = np.zeros(image, np.int32), np.zeros(image, np.int32)
    tw = np.zeros((image, cap), np.int32)
    mt, prec = C.c_int(0), C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    rc = lib.ap_resize_tap_tables(filt, n_in, n_out, image, p(tmin), p(tcnt), p(tw), cap, C.byref(mt), C.byref(prec))
    return rc, tmin, tcnt, tw, mt.value, prec.value


SIZES = [(224, 256, 224), (512, 256, 224), (256, 256, 224), (300, 256, 224), (224, 224, 224), (256, 224, 224), (512, 224, 224), (448, 224, 224),
         (1024, 224, 224), (235, 235, 224), (300, 235, 224), (257, 248, 224), (100, 224, 224)]


@pytest.mark.parametrize("filt", [0, 1, 2, 3])
@pytest.mark.parametrize("n_in,n_out,image", SIZES)
def test_tables_equal_the_library_coefficients(lib, filt, n_in, n_out, image):
    rc, tmin, tcnt, tw, max_taps, prec = _tables(lib, filt, n_in, n_out, image)
    assert rc == 0
    if filt in (0, 2):      # ATen: int16 weights, precision chosen over all outputs; transformers' crop offset (h - crop) // 2
        xmins, sizes, ws, want_prec = ra.aa_weights(n_in, n_out, "bicubic" if filt == 0 else "bilinear")
        off = (n_out - image) // 2
    else:                   # Pillow: 22-bit coefficients; torchvision's crop offset int(round((h - crop) / 2))
        xmins, sizes, ws = ra.pil_bilinear_weights(n_in, n_out, "bilinear" if filt == 1 else "bicubic")
        want_prec, off = ra.PIL_PRECISION_BITS, int(round((n_out - image) / 2.0))
    assert prec == want_prec and max_taps == int(max(sizes))
    for o in range(image):
        i = o + off
        assert tmin[o] == xmins[i] and tcnt[o] == sizes[i], (o, i)
        assert np.array_equal(tw[o, :sizes[i]], ws[i]) and not tw[o, sizes[i]:].any(), (o, i)


def test_capacity_and_arguments_are_checked(lib):
    rc, *_ , max_taps, _p = _tables(lib, 0, 1024, 224, 224, cap=4)
    assert rc != 0 and max_taps > 4                     # tells how much room is needed
    assert _tables(lib, 7, 224, 256, 224)[0] != 0       # unknown filter
    assert _tables(lib, 0, 224, 200, 224)[0] != 0       # crop larger than the resized image


@pytest.mark.parametrize("n_src,n_dst", [(512, 256), (257, 256), (320, 256), (384, 256), (683, 256), (768, 256), (1024, 256), (2048, 256),
                                         (448, 224), (300, 224), (1000, 224), (513, 512)])
def test_cv2_linear_taps_equal_the_opencv_arithmetic(lib, n_src, n_dst):
    """The taps of the cv2.resize the reference applies to reads larger than the patch (40x / 80x slides at 20x), against the
    restatement that tests/test_oracle_filter.py pins to cv2 itself."""
    from oracle.patch_filter import linear_taps

    taps, w = np.zeros(2 * n_dst, np.int32), np.zeros(2 * n_dst, np.int16)
    assert lib.ap_linear_tap_tables(n_src, n_dst, taps.ctypes.data_as(C.POINTER(C.c_int32)), w.ctypes.data_as(C.POINTER(C.c_int16))) == 0
    s, a0, a1 = linear_taps(n_src, n_dst)
    assert np.array_equal(taps[0::2], s) and np.array_equal(taps[1::2], s + 1)
    assert np.array_equal(w[0::2], a0) and np.array_equal(w[1::2], a1) and np.all(w[0::2].astype(int) + w[1::2] == 2048)
    assert s.min() >= 0 and s.max() + 1 <= n_src - 1      # down-scaling never leaves the source: no border rule is needed


def test_cv2_linear_taps_refuse_up_scaling(lib):
    taps, w = np.zeros(8, np.int32), np.zeros(8, np.int16)
    assert lib.ap_linear_tap_tables(200, 256, taps.ctypes.data_as(C.POINTER(C.c_int32)), w.ctypes.data_as(C.POINTER(C.c_int16))) != 0
