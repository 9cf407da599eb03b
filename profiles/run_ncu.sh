#!/bin/bash
# Run on the GPU box through gpurun (B200_PROFILING.md recipe).  Outputs land in gpurun_out/ and are copied to profiles/
# by hand after reading them.  Never a bench value: numbers printed under ncu are discarded.
set -x
export BENCH_QUICK=1
TAG=${1:-r01}
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
# 2. the named kernel of BASELINE configs[1] (fused demod+resample), full set, source view
ncu --set full --clock-control none --import-source on -k regex:rs_main -s 2 -c 2 -o gpurun_out/${TAG}_rs_main \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_rs_main.log 2>&1
# 3. the FFT pass kernel and the sync-search kernel
ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 8 -c 4 -o gpurun_out/${TAG}_fft_pass \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_fft_pass.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fs_sync$ -s 2 -c 1 -o gpurun_out/${TAG}_fs_sync \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_fs_sync.log 2>&1
# 4. the rest of the frame stage, one launch each
ncu --set full --clock-control none --import-source on -k 'regex:fs_collapse|fs_shift|fs_norm_lowpass|fs_minmax' -s 8 -c 4 -o gpurun_out/${TAG}_frame \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_frame.log 2>&1
ls -la gpurun_out/
