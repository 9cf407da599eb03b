mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gputests.log 2>&1; echo "exit $?" >> gpurun_out/r02d_gputests.log); tail -5 gpurun_out/r02d_gputests.log
for V in 0 1; do if [ $V = 1 ]; then export BENCH_NO_FRD_OVERLAP=1; fi; BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 > gpurun_out/r02d_quick_frdoverlap_off$V.json 2>/dev/null; done; unset BENCH_NO_FRD_OVERLAP
cat gpurun_out/r02d_quick_frdoverlap_off*.json | cut -c1-130
timeout 500 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -c 600 gpurun_out/r02d_bench.json; tail -3 gpurun_out/r02d_bench.err
