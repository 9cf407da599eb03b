"""Design study for the next round: the autocorrelation with both transforms at HALF size.

The capture is real and so is |X|/N, the input of the inverse transform.  A real sequence of length N packs into a complex
one of length N/2 (z[j] = x[2j] + i x[2j+1]); one N/2-point transform then yields the transforms E, O of the even and odd
samples (E = (Z + conj(Z mirrored))/2, O = (Z - conj(Z mirrored))/2i), and the reference's LAST radix-2 stage -- with ITS
perturbed angle (fft.c:161, tsdrgpu_fft_reference_eps) -- combines them: X[k] = E[k] + w^k O[k], X[k+N/2] = E[k] - w^k O[k].
The sub-transforms keep the reference's perturbed stage angles too (the transform is linear), so the ONLY approximation is the
mirror identity, which holds exactly for unperturbed stages and to ~pi*eps_l for the perturbed ones below the last.

This script measures that approximation in float64 against (a) a float64 model of the reference's transform and (b) the
compiled reference itself (float32), at the capture size of BASELINE configs[1] (N = 2^20).  CPU only, numpy only.

    python profiles/studies/real_input_autocorr_study.py [log2N]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def reference_eps(stages):
    from tempestsdr_b200 import _native
    eps = (C.c_double * stages)()
    _native.lib().tsdrgpu_fft_reference_eps(stages, 0, eps)
    return np.array(eps[:])


def bitrev(n):
    m = n.bit_length() - 1
    idx = np.arange(n)
    rev = np.zeros(n, dtype=np.int64)
    for b in range(m):
        rev |= ((idx >> b) & 1) << (m - 1 - b)
    return rev


def pfft(x, eps, inverse=False):
    """float64 model of fft_perform (fft.c:96-176): radix-2 DIT, stage l rotates by (pi/2^l)(1+eps_l); forward scales by 1/N."""
    n = x.size
    m = n.bit_length() - 1
    a = x.astype(np.complex128)[bitrev(n)]
    sign = 1.0 if inverse else -1.0
    for l in range(m):
        half = 1 << l
        w = np.exp(sign * 1j * np.pi * np.arange(half) / half * (1.0 + eps[l]))
        a = a.reshape(-1, 2, half)
        t = a[:, 1, :] * w
        a = np.stack([a[:, 0, :] + t, a[:, 0, :] - t], axis=1).reshape(-1)
    return a if inverse else a / n


def pfft_real_half(x, eps, inverse=False):
    """The same transform of a REAL x through one N/2-point transform + the reference's last stage."""
    n = x.size
    m = n.bit_length() - 1
    z = x[0::2] + 1j * x[1::2]
    zt = pfft(z, eps[: m - 1], inverse)
    if not inverse:
        zt = zt * (n // 2)                                 # undo the sub-transform's own 1/(N/2)
    mir = np.conj(np.roll(zt[::-1], 1))                    # conj(Z[(N/2 - k) mod N/2])
    e, o = 0.5 * (zt + mir), -0.5j * (zt - mir)
    half = n // 2
    sign = 1.0 if inverse else -1.0
    w = np.exp(sign * 1j * np.pi * np.arange(half) / half * (1.0 + eps[m - 1]))
    out = np.concatenate([e + w * o, e - w * o])
    return out if inverse else out / n


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << log2n
    eps = reference_eps(log2n)
    rng = np.random.default_rng(1)
    # a capture-like input: magnitudes of noisy video-like IQ
    x = np.abs(rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.float32)
    x[::1000] += 3.0
    xd = x.astype(np.float64)
    # model, full size
    X = pfft(xd, eps)
    R = np.abs(X)
    y_full = pfft(R, eps, inverse=True)
    # half-size route for both transforms
    Xh = pfft_real_half(xd, eps)
    Rh = np.abs(Xh)
    y_half = pfft_real_half(Rh, eps, inverse=True)
    peak = np.abs(y_full).max()
    print(f"N = 2^{log2n}; eps of the last three stages: {eps[-3:]}")
    print(f"forward : max |X_half - X_full| / max|X|      = {np.abs(Xh - X).max() / np.abs(X).max():.3e}")
    print(f"autocorr: max |y_half - y_full| / zero-lag peak = {np.abs(y_half - y_full).max() / peak:.3e}")
    lo, hi = int(25e6 / 87), int(25e6 / 55)
    if hi < n:
        print(f"          the same over the frame-lag window [{lo}, {hi}) = {np.abs(y_half[lo:hi] - y_full[lo:hi]).max() / peak:.3e}")
    try:
        from oracle import oracle as orc
        O = orc.best()
        want = O.autocorrelation(x)[: 2 * n].astype(np.float64)
        ref = want[0::2] + 1j * want[1::2]
        pk = np.abs(ref).max()
        print(f"compiled reference (float32) vs float64 model, full size : {np.abs(ref - y_full).max() / pk:.3e}  of the peak")
        print(f"compiled reference (float32) vs float64 model, half size : {np.abs(ref - y_half).max() / pk:.3e}  of the peak")
    except Exception as e:                                  # the study still stands on the model alone
        print("oracle not available:", e)


if __name__ == "__main__":
    main()
