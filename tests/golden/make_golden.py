"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref, built in place from /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
Every vector is the output of the reference's own compiled code on a seeded input that the test regenerates
from the same seed (inputs are stored too when they are small).  tests/test_golden.py checks both the C
restatement (CPU) and the CUDA path (GPU) against these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc      # noqa: E402
from tempestsdr_b200 import synth    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
R = orc.ref()


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name), **kw)
    print(name, {k: getattr(v, "shape", v) for k, v in kw.items()})


# --- demod + resample stream, cfg1 geometry scaled down (fs=800 kS/s, h=105, 60 Hz -> W=253, r=1.9924)
fs, h, fv = 800_000, 105, 60.0
w, _, _ = R.geometry(fs, h, fv)
block = int(0.1 * fs / fv)
iq = synth.video_like_iq(6 * block, fs, 2 * w // 2, h, fv, seed=1)
mag = R.am_demod(iq)
rs = R.resampler()
pix, states = [], []
for k in range(6):
    pix.append(rs.run(mag[k * block:(k + 1) * block], w * h * fv, fs))
    states.append(rs.state)
rs = R.resampler()
pix_nn = [rs.run(mag[k * block:(k + 1) * block], w * h * fv, fs, True) for k in range(6)]
save("demod_resample.npz", fs=fs, h=h, fv=fv, w=w, block=block, seed=1, mag=mag, pixels=np.concatenate(pix),
     counts=np.array([p.size for p in pix]), states=np.array(states), pixels_nn=np.concatenate(pix_nn))

# --- frame stage: 8 frames, GUI-default flags, 253x105
pp = R.postprocessor(fs, h, fv, autoshift=1, pll=0)
frames_in, frames_out, meta = [], [], []
for k in range(8):
    f = synth.video_like_frame(w, h, seed=100 + k, shift_x=31 + 5 * k, shift_y=11 + k)
    o, res = pp.run(f, w, h, 0.0, 0.1, 1, 0)
    frames_out.append(o)
    meta.append([res.x.dx, res.x.vx, res.x.curr_stripsize, res.y.dx, res.y.vx, res.y.curr_stripsize])
save("frame_stage_default.npz", w=w, h=h, seeds=np.arange(100, 108), out=np.stack(frames_out),
     meta=np.array(meta, dtype=np.int32))

# --- frame stage with motion blur, autogain after, lowpass after sync
pp = R.postprocessor(fs, h, fv, autoshift=1, pll=0)
frames_out, meta = [], []
for k in range(6):
    f = synth.video_like_frame(w, h, seed=200 + k, shift_x=40, shift_y=9)
    o, res = pp.run(f, w, h, 0.35, 0.1, 0, 1)
    frames_out.append(o)
    meta.append([res.x.dx, res.x.vx, res.x.curr_stripsize, res.y.dx, res.y.vx, res.y.curr_stripsize])
save("frame_stage_blur.npz", w=w, h=h, seeds=np.arange(200, 206), out=np.stack(frames_out),
     meta=np.array(meta, dtype=np.int32))

# --- FFT / autocorrelation
x = synth.noise_iq(4096, seed=5)
save("fft_4096.npz", x=x, fwd=R.fft(x, False), inv=R.fft(x, True))
cap = np.abs(synth.noise_iq(20_000, seed=6)[:20_000]).astype(np.float32)
save("autocorr_20000.npz", x=cap, ac=R.autocorrelation(cap))
fsd = 1_000_000
size = int(3.1 * fsd / 55.0)
det = R.framerate_detector()
plots = []
for k in range(2):
    c = R.am_demod(synth.video_like_iq(size, fsd, 300, 120, 55.5, seed=30 + k))
    (fo, fp), (lo, lp), calls = det.run(fsd, c)
    plots.append((fp, lp))
save("framerate_plots.npz", fs=fsd, size=size, seeds=np.array([30, 31]), frame_off=fo, line_off=lo,
     frame_plot=plots[-1][0], line_plot=plots[-1][1])

# --- superbandwidth stitch, 4 hops of 2^14
fss, fvs = 200_000, 50.0
sif = int(fss / fvs)
pairs = 5 * sif
base = synth.video_like_iq(pairs + 3000, fss, 100, 80, fvs, seed=9, snr_db=25)
lags = (0, 700, 33, 1999)
hops = [base[2 * l: 2 * (l + pairs)].copy() + synth.noise_iq(pairs, seed=100 + i, scale=0.01) for i, l in enumerate(lags)]
out, offs = R.superb_ondataready(hops, sif)
save("superb_4x16384.npz", fs=fss, fv=fvs, sif=sif, pairs=pairs, lags=np.array(lags), offsets=offs, out=out)
