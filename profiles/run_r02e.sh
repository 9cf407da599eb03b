mkdir -p gpurun_out
BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 > gpurun_out/r02e_quick_prio_default.json 2>/dev/null
BENCH_MAIN_STREAM_PRIO=-1 BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 > gpurun_out/r02e_quick_prio_main_high.json 2>/dev/null
BENCH_MAIN_STREAM_PRIO=-1 TSDRGPU_FRD_PRIO=high BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 > gpurun_out/r02e_quick_prio_all_high.json 2>/dev/null
for f in default main_high all_high; do echo $f; cut -c1-110 gpurun_out/r02e_quick_prio_$f.json; done
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "superband or framerate or autocorr" 2>&1 | tail -3)
