#!/bin/bash
# Targeted captures (one launch list + named kernels); same recipe as run_ncu.sh.
set -x
export BENCH_QUICK=1
TAG=${1:-r01}
shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 2 -c 2 -o gpurun_out/${TAG}_${K//[^a-zA-Z0-9_]/} \
      python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_${K//[^a-zA-Z0-9_]/}.log 2>&1
done
ls -la gpurun_out/ | tail -20
