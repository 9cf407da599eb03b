#!/usr/bin/env python3
"""How long does rank 0's one-GPU reference stitch take at H hops of 2^21 (first call, later calls)?  bench.py runs it on rank 0
while the other ranks already wait inside the sharded stitch's flag kernels (4 s timeout).   python profiles/studies/one_gpu_stitch_timing.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tempestsdr_b200 import api  # noqa: E402
from tempestsdr_b200.api import PostProcessFlags  # noqa: E402

gpu = api.Context(0)
FS, FV, HEIGHT = bench.FS, bench.FV, bench.HEIGHT
sif = int(FS / FV)
hop_pairs = 10 * sif
out = {}
t0 = time.perf_counter()
src = torch.from_numpy(bench.make_iq(hop_pairs + 16 * 1000, seed=4242)).cuda()
out["make_iq_s"] = time.perf_counter() - t0
for H in (2, 4, 8):
    offs = [0] + [131 + 977 * q for q in range(1, H)]
    hops = [src[2 * offs[q]: 2 * (offs[q] + hop_pairs)].contiguous() for q in range(H)]
    times = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        one_iq, one_offs = gpu.superb_stitch(hops, sif)
        torch.cuda.synchronize(); times.append(round(1e3 * (time.perf_counter() - t0), 2))
    row = {"superb_stitch_ms": times}
    t0 = time.perf_counter()
    n_fft = gpu.fft_getrealsize(hop_pairs)
    wH = int(2 * (H * FS / (FV * HEIGHT)))
    blockH = int(0.1 * H * FS / FV)
    nblk = (H * n_fft) // blockH
    rs, pp = gpu.resampler(), gpu.post_processor()
    upH = float(wH * HEIGHT) * FV
    pix = torch.empty(int(rs.plan((blockH, nblk), upH, float(H * FS))) + 1024, dtype=torch.float32, device="cuda")
    nH = wH * HEIGHT
    frames_out = torch.empty(((pix.numel() // nH) + 1) * nH, dtype=torch.float32, device="cuda")
    mag = gpu.am_demod(one_iq)
    torch.cuda.synchronize(); row["setup_ms"] = round(1e3 * (time.perf_counter() - t0), 2)
    rt = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        px = rs.process(mag, (blockH, nblk), upH, float(H * FS), in_is_iq=False, out=pix)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        nf = px.numel() // nH
        pp.process(px[: nf * nH], wH, HEIGHT, 0.0, 0.1, PostProcessFlags(autoshift=True, lowpass_before_sync=True, superresolution=True),
                   out=frames_out[: nf * nH], want_results=False)
        torch.cuda.synchronize(); rt.append([round(1e3 * (t1 - t0), 2), round(1e3 * (time.perf_counter() - t1), 2)])
    row["resample_ms__framestage_ms"] = rt; row["frames"] = int(nf); row["frame"] = [wH, HEIGHT]
    out[f"H={H}"] = row
print(json.dumps(out, indent=1))
