"""ctypes binding of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two libraries expose the same flat C API (oracle/tsdr_oracle.h):

* ``port()``  -> oracle/libtsdr_oracle.so           my C restatement (prefix ``orc_``), always available;
* ``ref()``   -> oracle/_ref/libtsdr_refharness.so   the real reference compiled in place (prefix ``refh_``),
                 available where it was built (this container; the prebuilt .so travels to the GPU box).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` leg may import this
module.  The product (tempestsdr_b200/) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libtsdr_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libtsdr_refharness.so")
REF_LIB_SO = os.path.join(_HERE, "_ref", "libTSDRLibrary.so")
REF_RAWFILE_SO = os.path.join(_HERE, "_ref", "libTSDRPlugin_RawFile.so")
REF_RAWFILE_NOPACE_SO = os.path.join(_HERE, "_ref", "libTSDRPlugin_RawFile_nopace.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Autogain(C.Structure):
    _fields_ = [("lastmax", C.c_float), ("lastmin", C.c_float), ("snr", C.c_float)]


class Sweetspot(C.Structure):
    _fields_ = [("dx", C.c_int), ("vx", C.c_int), ("absvx", C.c_int), ("curr_stripsize", C.c_int)]

    def astuple(self):
        return (self.dx, self.vx, self.absvx, self.curr_stripsize)


class Dropcomp(C.Structure):
    _fields_ = [("difference", C.c_int64)]


class PPConfig(C.Structure):
    _fields_ = [("samplerate", C.c_uint32), ("height", C.c_int), ("refreshrate", C.c_double),
                ("autoshift", C.c_int), ("pll", C.c_int), ("superres", C.c_int)]


class PPResult(C.Structure):
    _fields_ = [("x", Sweetspot), ("y", Sweetspot), ("avg_speed", C.c_double), ("pll_state", C.c_int),
                ("lastmax", C.c_float), ("lastmin", C.c_float), ("snr", C.c_float),
                ("refreshrate_after", C.c_double), ("width_after", C.c_int),
                ("pll_callback_fired", C.c_int), ("autogain_callback_fired", C.c_int),
                ("autogain_cb_min", C.c_double), ("autogain_cb_max", C.c_double)]


def build_port() -> None:
    """(Re)build the C restatement (and the reference binaries when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


class Oracle:
    """One of the two oracle libraries behind a numpy-friendly facade."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.kind = kind
        self.path = path
        self._lib = C.CDLL(path)
        self._p = prefix
        L = self._fn
        L("am_demod", None, [f32p, C.c_int, f32p])
        L("resample_new", C.c_void_p, [])
        L("resample_free", None, [C.c_void_p])
        L("resample_get", None, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)])
        L("resample_set", None, [C.c_void_p, C.c_double, C.c_double])
        L("resample_run", C.c_uint32, [C.c_void_p, f32p, C.c_uint32, C.c_double, C.c_double, C.c_int, f32p, C.c_uint32])
        L("dropcomp_shift_with", None, [C.POINTER(Dropcomp), C.c_uint32, C.c_int64])
        L("dropcomp_will_drop_all", C.c_int, [C.POINTER(Dropcomp), C.c_uint32, C.c_uint32])
        L("dropcomp_add", C.c_uint32, [C.POINTER(Dropcomp), C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)])
        L("geometry", None, [C.c_uint32, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)])
        L("autogain", None, [C.POINTER(Autogain), C.c_int, f32p, f32p, C.c_float])
        L("timelowpass", None, [C.c_float, C.c_int, f32p, f32p])
        L("average_v_h", None, [C.c_int, C.c_int, f32p, f32p, f32p])
        L("gaussianblur", None, [f32p, C.c_int])
        L("findbestfit", None, [f32p, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)])
        L("findthesweetspot", None, [C.POINTER(Sweetspot), f32p, C.c_int, C.c_int, C.c_double])
        L("pp_new", C.c_void_p, [])
        L("pp_free", None, [C.c_void_p])
        L("pp_config", None, [C.c_void_p, C.POINTER(PPConfig)])
        L("pp_run", C.c_int, [C.c_void_p, f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, f32p, C.POINTER(PPResult)])
        L("pixels_argb", None, [f32p, C.c_int, C.c_int, C.c_void_p, i32p])
        L("fft_getrealsize", C.c_uint32, [C.c_uint32])
        L("fft", None, [f32p, C.c_uint32, C.c_int])
        L("autocorrelation", None, [f32p, f32p, C.c_uint32])
        L("crosscorrelation", None, [f32p, f32p, C.c_uint32])
        L("accumulate", None, [f64p, C.c_uint64, f32p, C.c_int, C.c_int])
        L("frd_new", C.c_void_p, [])
        L("frd_free", None, [C.c_void_p])
        L("frd_run", C.c_int, [C.c_void_p, C.c_uint32, f32p, C.c_int, f64p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                               f64p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)])
        L("complex_to_abs_diff", None, [f32p, C.c_int])
        L("superb_bestfit", C.c_int, [f32p, f32p, C.c_int, C.c_int])
        L("superb_ondataready", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, f32p, i32p])
        if kind == "port":
            L("resample_last_emitted", C.c_uint32, [C.c_void_p])
            L("framerate_windows", None, [C.c_uint32] + [C.POINTER(C.c_int)] * 4)
            L("framerate_capture_size", C.c_uint32, [C.c_uint32])

    def _fn(self, name, restype, argtypes):
        f = getattr(self._lib, self._p + name)
        f.restype = restype
        f.argtypes = argtypes
        setattr(self, "_" + name, f)

    # ---- a2
    def am_demod(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        out = np.empty(iq.size // 2, dtype=np.float32)
        self._am_demod(iq, iq.size // 2, out)
        return out

    # ---- a6
    def resampler(self):
        return _Resampler(self)

    # ---- a3
    def dropcomp_shift_with(self, diff: int, block: int, syncoffset: int) -> int:
        d = Dropcomp(diff)
        self._dropcomp_shift_with(C.byref(d), block, syncoffset)
        return d.difference

    def dropcomp_will_drop_all(self, diff: int, size: int, block: int) -> int:
        d = Dropcomp(diff)
        return self._dropcomp_will_drop_all(C.byref(d), size, block)

    def dropcomp_add(self, diff: int, size: int, block: int, ring_accepts: bool):
        d = Dropcomp(diff)
        skip = C.c_uint32(0)
        fwd = self._dropcomp_add(C.byref(d), size, block, int(ring_accepts), C.byref(skip))
        return d.difference, fwd, skip.value

    # ---- a4
    def geometry(self, samplerate: int, height: int, refreshrate: float):
        w = C.c_int(0); pr = C.c_double(0); pt = C.c_double(0)
        self._geometry(samplerate, height, refreshrate, C.byref(w), C.byref(pr), C.byref(pt))
        return w.value, pr.value, pt.value

    # ---- a8-a10
    def autogain(self, state: Autogain, frame: np.ndarray, norm: float) -> np.ndarray:
        frame = np.ascontiguousarray(frame, dtype=np.float32)
        out = np.empty_like(frame)
        self._autogain(C.byref(state), frame.size, frame, out, norm)
        return out

    def timelowpass(self, coeff: float, frame: np.ndarray, screen: np.ndarray) -> None:
        self._timelowpass(coeff, frame.size, np.ascontiguousarray(frame, dtype=np.float32), screen)

    def average_v_h(self, frame: np.ndarray, w: int, h: int):
        wb = np.empty(w, dtype=np.float32); hb = np.empty(h, dtype=np.float32)
        self._average_v_h(w, h, np.ascontiguousarray(frame, dtype=np.float32), wb, hb)
        return wb, hb

    # ---- a12-a14
    def gaussianblur(self, strip: np.ndarray) -> np.ndarray:
        s = np.array(strip, dtype=np.float32, copy=True)
        self._gaussianblur(s, s.size)
        return s

    def findbestfit(self, strip: np.ndarray, totalsum: float, stripsize: int):
        bf = C.c_double(0); bi = C.c_int(0)
        self._findbestfit(np.ascontiguousarray(strip, dtype=np.float32), strip.size, totalsum, stripsize, C.byref(bf), C.byref(bi))
        return bf.value, bi.value

    def findthesweetspot(self, state: Sweetspot, strip: np.ndarray, minsize: int, lowpass: float) -> np.ndarray:
        s = np.array(strip, dtype=np.float32, copy=True)
        self._findthesweetspot(C.byref(state), s, s.size, minsize, lowpass)
        return s

    # ---- a7
    def postprocessor(self, samplerate: int, height: int, refreshrate: float, autoshift=1, pll=0, superres=0):
        return _PostProcessor(self, PPConfig(samplerate, height, refreshrate, autoshift, pll, superres))

    # ---- a16
    def pixels_argb(self, frame: np.ndarray, inverted: bool = False) -> np.ndarray:
        frame = np.ascontiguousarray(frame, dtype=np.float32)
        out = np.zeros(frame.size, dtype=np.int32)
        self._pixels_argb(frame, frame.size, int(inverted), None, out)
        return out

    # ---- a19, a20
    def fft_getrealsize(self, n: int) -> int:
        return self._fft_getrealsize(n)

    def fft(self, iq: np.ndarray, inverse: bool) -> np.ndarray:
        d = np.array(iq, dtype=np.float32, copy=True)
        self._fft(d, d.size // 2, int(inverse))
        return d

    def autocorrelation(self, real: np.ndarray) -> np.ndarray:
        real = np.ascontiguousarray(real, dtype=np.float32)
        ans = np.empty(2 * real.size, dtype=np.float32)
        self._autocorrelation(ans, real, real.size)
        return ans

    def crosscorrelation(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.array(a, dtype=np.float32, copy=True); b = np.array(b, dtype=np.float32, copy=True)
        self._crosscorrelation(a, b, a.size // 2)
        return a

    # ---- a18
    def accumulate(self, out: np.ndarray, calls: int, ac: np.ndarray, startid: int, length: int) -> None:
        self._accumulate(out, calls, np.ascontiguousarray(ac, dtype=np.float32), startid, length)

    def framerate_detector(self):
        return _FrameRateDetector(self)

    def framerate_windows(self, samplerate: int):
        v = [C.c_int(0) for _ in range(4)]
        self._framerate_windows(samplerate, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def framerate_capture_size(self, samplerate: int) -> int:
        return self._framerate_capture_size(samplerate)

    # ---- a22
    def complex_to_abs_diff(self, iq: np.ndarray) -> np.ndarray:
        d = np.array(iq, dtype=np.float32, copy=True)
        self._complex_to_abs_diff(d, d.size)
        return d

    def superb_bestfit(self, a: np.ndarray, b: np.ndarray, samples_in_frame: int) -> int:
        a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
        return self._superb_bestfit(a, b, a.size, samples_in_frame)

    def superb_ondataready(self, hops, samples_in_frame: int):
        """hops: list of interleaved IQ float32 arrays of equal length.  Returns (stitched IQ, offsets)."""
        bufs = [np.array(h, dtype=np.float32, copy=True) for h in hops]
        pairs = bufs[0].size // 2
        n = self.fft_getrealsize(pairs)
        out = np.zeros(len(bufs) * n * 2, dtype=np.float32)
        offs = np.zeros(len(bufs), dtype=np.int32)
        arr = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        total = self._superb_ondataready(arr, len(bufs), pairs, samples_in_frame, out, offs)
        return out[: 2 * total], offs


class _Resampler:
    def __init__(self, o: Oracle):
        self._o = o
        self._h = o._resample_new()

    def __del__(self):
        try:
            self._o._resample_free(self._h)
        except Exception:
            pass

    @property
    def state(self):
        c = C.c_double(0); f = C.c_double(0)
        self._o._resample_get(self._h, C.byref(c), C.byref(f))
        return c.value, f.value

    @state.setter
    def state(self, cf):
        self._o._resample_set(self._h, cf[0], cf[1])

    @property
    def last_emitted(self) -> int:
        return self._o._resample_last_emitted(self._h)

    def run(self, block: np.ndarray, upsample_by: float, downsample_by: float, nearest: bool = False) -> np.ndarray:
        block = np.ascontiguousarray(block, dtype=np.float32)
        cap = int(block.size * (upsample_by / downsample_by)) + 16
        out = np.empty(cap, dtype=np.float32)
        n = self._o._resample_run(self._h, block, block.size, upsample_by, downsample_by, int(nearest), out, cap)
        return out[:n].copy()


class _PostProcessor:
    def __init__(self, o: Oracle, cfg: PPConfig):
        self._o = o
        self._h = o._pp_new()
        self.cfg = cfg
        o._pp_config(self._h, C.byref(cfg))

    def __del__(self):
        try:
            self._o._pp_free(self._h)
        except Exception:
            pass

    def run(self, frame: np.ndarray, w: int, h: int, motionblur=0.0, lowpasscoeff=0.1, lowpass_before_sync=1,
            autogain_after=0):
        frame = np.ascontiguousarray(frame, dtype=np.float32)
        assert frame.size == w * h
        out = np.empty(w * h, dtype=np.float32)
        res = PPResult()
        rc = self._o._pp_run(self._h, frame, w, h, motionblur, lowpasscoeff, lowpass_before_sync, autogain_after, out, C.byref(res))
        assert rc == 0
        return out, res


class _FrameRateDetector:
    def __init__(self, o: Oracle):
        self._o = o
        self._h = o._frd_new()

    def __del__(self):
        try:
            self._o._frd_free(self._h)
        except Exception:
            pass

    def run(self, samplerate: int, capture: np.ndarray):
        capture = np.ascontiguousarray(capture, dtype=np.float32)
        fcap = int(samplerate / 55) + 8
        lcap = int(samplerate / (590 * 55)) + 8
        fp = np.zeros(fcap); lp = np.zeros(lcap)
        fo = C.c_int(0); fl = C.c_int(0); lo = C.c_int(0); ll = C.c_int(0); calls = C.c_uint64(0)
        rc = self._o._frd_run(self._h, samplerate, capture, capture.size, fp, fcap, C.byref(fo), C.byref(fl),
                              lp, lcap, C.byref(lo), C.byref(ll), C.byref(calls))
        assert rc == 0
        return (fo.value, fp[: fl.value].copy()), (lo.value, lp[: ll.value].copy()), calls.value


_cache = {}


def port() -> Oracle:
    if "port" not in _cache:
        if not os.path.exists(PORT_SO):
            build_port()
        _cache["port"] = Oracle(PORT_SO, "orc_", "port")
    return _cache["port"]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> Oracle:
    if "ref" not in _cache:
        if not have_ref():
            raise FileNotFoundError("oracle/_ref not built (needs /root/reference): run `make -C oracle ref`")
        _cache["ref"] = Oracle(REF_SO, "refh_", "reference")
    return _cache["ref"]


def best() -> Oracle:
    """The real reference when its binary is present, else the pinned restatement."""
    return ref() if have_ref() else port()
