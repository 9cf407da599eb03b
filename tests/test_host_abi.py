"""CPU-only checks of the product's host side: the C-ABI library loads without a GPU, exports every symbol the
header declares, fails loudly (no fallback) when no device exists, and its host-side scalar planning agrees with the
oracle bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from tempestsdr_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(path):
    txt = open(path).read()
    return sorted(set(re.findall(r"\b(tsdrgpu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    declared = header_symbols(os.path.join(ROOT, "include", "tsdrgpu.h"))
    assert len(declared) > 40
    for s in declared:
        assert hasattr(lib, s), f"libtsdrgpu.so does not export {s}"
    assert sorted(_native.DECLARED_SYMBOLS) == declared, "python binding table out of sync with the header"


def test_no_silent_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _native.lib()
    assert lib.tsdrgpu_device_count() == 0
    h = C.c_void_p()
    rc = lib.tsdrgpu_create(C.byref(h), 0)
    assert rc == -1 and not h.value                      # TSDRGPU_ENODEVICE
    assert b"no CPU fallback" in lib.tsdrgpu_last_error(None)
    from tempestsdr_b200 import api
    with pytest.raises(api.TsdrGpuError):
        api.Context(0)


@pytest.mark.parametrize("fs,h,fv", [(8_000_000, 525, 60.0), (25_000_000, 1125, 60.0), (50_000_000, 1125, 60.0),
                                     (100_000_000, 2250, 60.0), (1_000_000, 100, 50.0), (2_400_000, 313, 59.94)])
def test_host_geometry_and_resample_plan(fs, h, fv):
    lib = _native.lib()
    P = orc.port()
    w, pr, pt = C.c_int(0), C.c_double(0), C.c_double(0)
    lib.tsdrgpu_geometry(fs, h, fv, C.byref(w), C.byref(pr), C.byref(pt))
    assert (w.value, pr.value, pt.value) == P.geometry(fs, h, fv)
    # the phase recurrence over 200 decimator blocks equals the oracle's resampler state block by block
    block = int(0.1 * fs / fv)
    up = w.value * h * fv
    rs = P.resampler()
    off = C.c_double(0.0)
    x = np.zeros(block, np.float32)
    for k in range(200):
        n = lib.tsdrgpu_plan_resample(C.byref(off), None, block, 1, up, float(fs), None)
        assert n == rs.run(x, up, float(fs)).size
        assert off.value == rs.state[1]
    # and in one call of 200 blocks
    off2 = C.c_double(0.0)
    blocks = (C.c_byte * (40 * 200))()
    total = lib.tsdrgpu_plan_resample(C.byref(off2), None, block, 200, up, float(fs), blocks)
    assert off2.value == off.value and total > 0
    assert lib.tsdrgpu_plan_resample(C.byref(C.c_double(0.0)), None, 1, 1, 0.1, 1.0, None) == 2**64 - 1


def test_gauss_taps_match_oracle():
    lib = _native.lib()
    taps = (C.c_float * 5)()
    lib.tsdrgpu_gauss_taps(C.byref(taps))
    imp = np.zeros(11, np.float32); imp[5] = 1.0
    blurred = orc.port().gaussianblur(imp)
    assert np.array_equal(np.array(taps[:], np.float32)[::-1].view(np.uint32), blurred[3:8].view(np.uint32))


def test_detect_videomode_matches_the_gui_arithmetic():
    """tsdrgpu_detect_videomode (host-only) against a restatement of PlotVisualizer.java:203-236 + Main.java:1301-1303,1346-1350."""
    import ctypes as C
    import numpy as np
    from tempestsdr_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(8)
    for trial in range(50):
        fs = int(rng.integers(2_000_000, 60_000_000))
        foff, flen = fs // 87, fs // 55 - fs // 87                     # frameratedetector.c:20-25 windows
        loff, llen = int(fs / (1500 * 87.0)), int(fs / (590 * 55.0)) - int(fs / (1500 * 87.0))
        fp = rng.random(flen); lp = rng.random(max(llen, 1))
        if trial % 3 == 0:                                             # ties: the FIRST maximum wins
            fp[[5, 9]] = 2.0; lp[[1, min(3, lp.size - 1)]] = 2.0
        fi, li = int(np.argmax(fp)), int(np.argmax(lp))
        want_fps = float(fs) / float(foff + fi)
        want_h = int(np.floor((foff + fi) / float(loff + li) + 0.5))
        fps, h, a, b = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        rc = lib.tsdrgpu_detect_videomode(fp.ctypes.data_as(C.POINTER(C.c_double)), foff, fp.size, lp.ctypes.data_as(C.POINTER(C.c_double)), loff, lp.size,
                                          fs, C.byref(fps), C.byref(h), C.byref(a), C.byref(b))
        assert rc == 0 and (a.value, b.value) == (fi, li) and fps.value == want_fps and h.value == want_h
    assert lib.tsdrgpu_detect_videomode(None, 0, 0, None, 0, 0, 1, None, None, None, None) != 0


def test_resample_plan_random_sweep_against_the_oracle():
    """Host planning (output count per block, carried phase) for random geometries, ratios below and above 1 and ragged
    block sizes equals the oracle's resampler block by block: the data-independent half of dsp_resample_process
    (dsp.c:256-307) that the CUDA path trusts the host for."""
    import ctypes as C
    import numpy as np
    from tempestsdr_b200 import _native
    lib = _native.lib()
    P = orc.port()
    rng = np.random.default_rng(2024)
    for trial in range(40):
        down = float(rng.integers(1_000_000, 60_000_000))
        ratio = float(rng.choice([0.37, 0.5, 0.999, 1.0, 1.25, 1.998, 2.0, 3.3, 7.5])) * (1.0 + (rng.random() - 0.5) * 1e-3)
        up = down * ratio
        sizes = rng.integers(max(8, int(4 / ratio) + 4), 5000, size=12).astype(np.uint32)
        rs = P.resampler()
        off = C.c_double(0.0)
        total = 0
        for s in sizes:
            one = np.array([s], dtype=np.uint32)
            n = lib.tsdrgpu_plan_resample(C.byref(off), one.ctypes.data_as(C.c_void_p), 0, 1, up, down, None)
            out = rs.run(np.zeros(int(s), np.float32), up, down)
            assert n == out.size, (trial, int(s), ratio)
            assert off.value == rs.state[1], (trial, int(s), ratio)
            total += n
        off2 = C.c_double(0.0)
        assert lib.tsdrgpu_plan_resample(C.byref(off2), sizes.ctypes.data_as(C.c_void_p), 0, len(sizes), up, down, None) == total
        assert off2.value == off.value
