#!/bin/bash
# Round-2 captures (recipe of /opt/skills/guides/B200_PROFILING.md): one launch list of the bench command + ncu --set full of the
# named kernels.  Numbers printed by runs under ncu are never used as bench values.   usage: profiles/run_ncu_r02.sh <tag> <kernel regex>...
export BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=1
TAG=${1:-r02}
shift
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
for K in "$@"; do
  N=${K//[^a-zA-Z0-9_]/}
  # skip the warm-up launches of the kernel (3 warm-up batches: 12 FFT passes, 3 of every once-per-batch kernel), capture the next ones
  if [[ $K == fft* ]]; then S=12; C=4; else S=3; C=2; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K} -s $S -c $C -o gpurun_out/${TAG}_${N} \
      python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_${N}.log 2>&1
done
ls -la gpurun_out/${TAG}_* | tail -20
