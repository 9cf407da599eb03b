"""All H ranks of a superbandwidth group on ONE device (see test_superbandwidth_one_hop_per_rank_on_one_device).

    python -m tests.sbm_one_device H

Against superb_ondataready + am_demod of the reference (superbandwidth.c:121-152, TSDRLibrary.c:244-262): lags exact, magnitudes
to 1e-5 of the peak (float32 transforms in a different summation order; north_star's bound for float intermediates), two
stitches in a row (epoch-valued flags)."""
import sys

import numpy as np
import torch

from oracle import oracle as orc
from tempestsdr_b200 import api, superband, synth


def run(H: int, full: bool = False) -> None:
    O = orc.best()
    # `full`: BASELINE configs[3]'s own hop size -- 10 frames of 25 MS/s IQ, N = 2^21 per hop, an H x 2^21-point inverse whose last
    # stages carry the reference's largest angle errors (2.6e-5 relative at stage 22): the sharded path has to follow them
    fs, fv = (25_000_000, 60.0) if full else (400_000, 50.0)
    sif = int(fs / fv)
    pairs = 10 * sif
    base = synth.video_like_iq(pairs + 9000, fs, 2576 if full else 200, 1125 if full else 160, fv, seed=19, snr_db=25)
    offsets = [0, 1234, 77, 3999, 512, 6001, 2500, 4242][:H]
    ctxs = [api.Context(0) for _ in range(H)]
    streams = [torch.cuda.Stream() for _ in range(H)]
    groups = superband.SuperbGroup.local(ctxs, pairs)
    for rnd in range(1 if full else 2):
        hops = [base[2 * l: 2 * (l + pairs)].copy() + synth.noise_iq(pairs, seed=100 * rnd + i, scale=0.01) for i, l in enumerate(offsets)]
        want_iq, offs = O.superb_ondataready(hops, sif)
        want = O.am_demod(want_iq)
        d_hops = [torch.from_numpy(h).cuda() for h in hops]
        torch.cuda.synchronize()
        out = None
        for r in range(H):
            with torch.cuda.stream(streams[r]):
                # round 0: every rank holds its own copy of hop 0 (the alignment reference); round 1: rank 0's difference spectrum is pulled
                # round 1 also leaves the stream in the root's window (no copy out) and reads it from there
                res = groups[r].stitch(d_hops[r], sif, hop0=d_hops[0] if rnd == 0 else None, in_place=(rnd == 1))
                out = res if res is not None else out
        torch.cuda.synchronize()
        lags = groups[0].lags()
        assert [2 * l for l in lags] == [int(o) for o in offs], (rnd, lags, list(offs))
        assert groups[H - 1].lags() == lags
        got = out.cpu().numpy()
        assert got.shape == want.shape, (got.shape, want.shape)
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        print(f"H={H} round {rnd}: lags {lags}, max|err|/peak = {err:.3g}")
        assert err <= 1e-5, err
    print("sbm ok")


if __name__ == "__main__":
    run(int(sys.argv[1]), full=len(sys.argv) > 2 and sys.argv[2] == "full")
