#!/usr/bin/env python3
"""Where does an FFT pass spend its time?  The batched cfg2 autocorrelation (19 captures of 2^20 samples, half-size transforms)
timed four ways through TSDRGPU_FFT_DBG (csrc/fft.cu, timing experiments only -- the results of dbg runs are garbage):
    0  the real thing        1  no global loads in the passes        2  no global stores        3  neither (issue/shared-memory floor)
Run on the GPU box:  python profiles/studies/fft_floor_experiment.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tempestsdr_b200 import api  # noqa: E402

gpu = api.Context(0)
lib = gpu._lib
out = {}
for logn, batch in ((20, 19), (20, 64), (18, 64), (22, 16)):
    n = 1 << logn
    x = torch.rand(batch * n, device="cuda") + 0.25
    ans = torch.empty(batch * 2 * n, device="cuda")
    row = {}
    for dbg in (0, 1, 2, 3):
        os.environ["TSDRGPU_FFT_DBG"] = str(dbg)
        run = lambda: gpu.chk(lib.tsdrgpu_autocorrelation_batch(gpu._h, gpu.stream, ans.data_ptr(), x.data_ptr(), n, batch, n))
        for _ in range(3):
            run()
        gpu.chk(lib.tsdrgpu_profile_enable(gpu._h, 1))
        import bench  # noqa: E402  (collect_profile)
        bench.collect_profile(gpu)
        reps = 10
        for _ in range(reps):
            run()
        torch.cuda.synchronize()
        prof = bench.collect_profile(gpu)
        gpu.chk(lib.tsdrgpu_profile_enable(gpu._h, 0))
        row[f"dbg{dbg}"] = {k: round(1e3 * t / c, 2) for k, (t, c) in prof.items()}      # us per launch
    out[f"2^{logn} x{batch}"] = row
os.environ.pop("TSDRGPU_FFT_DBG", None)
print(json.dumps(out, indent=1))
