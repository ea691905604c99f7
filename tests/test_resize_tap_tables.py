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
