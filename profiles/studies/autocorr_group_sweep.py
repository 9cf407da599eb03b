"""How many captures should share the autocorrelation's work buffers?  (csrc/fft.cu autocorr_group)

Runs the frame-rate detector's batched autocorrelation (19 captures of the cfg2 size, as in one bench batch) with the group
size forced to several values and prints the device time per capture.  GPU only:  python profiles/studies/autocorr_group_sweep.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tempestsdr_b200 import api  # noqa: E402

FS = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
gpu = api.Context(0)
cap = api.FrameRateDetector.capture_size(FS)
batch = 19
x = torch.rand(batch * cap, device="cuda") + 0.25
flush = torch.empty(64 << 20, device="cuda")
for g in (0, 1, 2, 3, 4, 6, 8, 10, 12, 19):
    os.environ["TSDRGPU_AUTOCORR_GROUP"] = str(g)
    det = gpu.framerate_detector()
    for _ in range(3):
        det.run_batch(FS, x, cap, batch, cap)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); det.run_batch(FS, x, cap, batch, cap); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"group {g:2d} (0 = whole batch): {1e3 * ts[len(ts) // 2] / batch:7.2f} us per capture (median of 10, L2 flushed before each batch)")
