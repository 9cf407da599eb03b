mkdir -p gpurun_out
timeout 200 python profiles/studies/one_gpu_stitch_timing.py > gpurun_out/r02b_one_gpu_stitch.json 2> gpurun_out/r02b_one_gpu_stitch.err
timeout 200 python profiles/studies/fft_floor_experiment.py > gpurun_out/r02b_fft_floor.json 2> gpurun_out/r02b_fft_floor.err
for L in 0 8 16 32; do BENCH_L2_FRAMES=$L BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 > gpurun_out/r02b_quick_l2_$L.json 2>/dev/null; done
BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:rs_main4 -s 2 -c 2 -o gpurun_out/r02b_rs_main4 python bench.py --steps 1 --warmup 3 > gpurun_out/r02b_rs_main4.log 2>&1
timeout 500 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
cat gpurun_out/r02b_one_gpu_stitch.json; cat gpurun_out/r02b_fft_floor.json | head -80; cat gpurun_out/r02b_quick_l2_*.json; tail -3 gpurun_out/r02b_rs_main4.log; tail -c 1500 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
