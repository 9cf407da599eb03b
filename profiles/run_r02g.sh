mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gputests.log 2>&1; echo "exit $?" >> gpurun_out/r02g_gputests.log); tail -5 gpurun_out/r02g_gputests.log
for S in cfg5 cfg1; do BENCH_SHAPE=$S BENCH_NO_SWEEP=1 BENCH_NO_INT8=1 BENCH_NO_SHAPES=1 BENCH_E2E_SECONDS=0.5 BENCH_BATCHES_PER_STEP=8 timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench_$S.json 2> gpurun_out/r02g_bench_$S.err; python profiles/summarise.py gpurun_out/r02g_bench_$S.json | head -20; tail -2 gpurun_out/r02g_bench_$S.err; done
BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 2>/dev/null | cut -c1-120
