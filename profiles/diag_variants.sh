export BENCH_QUICK=1
echo "default:"; python bench.py --steps 10 --warmup 3
echo "no overlap:"; BENCH_NO_OVERLAP=1 python bench.py --steps 10 --warmup 3
echo "no fuse:"; TSDRGPU_NO_FUSE=1 python bench.py --steps 10 --warmup 3
echo "tables:"; TSDRGPU_FFT_TWIDDLE_TABLES=1 python bench.py --steps 10 --warmup 3
