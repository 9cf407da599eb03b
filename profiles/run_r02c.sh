mkdir -p gpurun_out
for M in 3 4 5 6; do TSDRGPU_RS_MINB=$M BENCH_QUICK=1 BENCH_BATCHES_PER_STEP=8 timeout 120 python bench.py --steps 6 --warmup 3 > gpurun_out/r02c_quick_minb_$M.json 2>/dev/null; done
for M in 3 4 5 6; do TSDRGPU_RS_MINB=$M python - <<'P' > gpurun_out/r02c_rs_minb_$M.txt 2>&1
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from tempestsdr_b200 import api
gpu = api.Context(0)
w = bench.geometry(); block = int(0.1 * bench.FS / bench.FV); nb = 640
iq = torch.from_numpy(bench.make_iq(block * nb, seed=5)).cuda()
rs = gpu.resampler(); up = float(w * bench.HEIGHT) * bench.FV
pix = torch.empty(int(rs.plan((block, nb), up, float(bench.FS))) + 1024, device="cuda"); mag = torch.empty(block * nb, device="cuda")
for _ in range(3): rs.process(iq, (block, nb), up, float(bench.FS), in_is_iq=True, out=pix, mag_out=mag)
gpu.chk(gpu._lib.tsdrgpu_profile_enable(gpu._h, 1)); bench.collect_profile(gpu)
for _ in range(10): rs.process(iq, (block, nb), up, float(bench.FS), in_is_iq=True, out=pix, mag_out=mag)
torch.cuda.synchronize()
print({k: round(1e3 * t / c, 2) for k, (t, c) in bench.collect_profile(gpu).items() if c})
P
done
cat gpurun_out/r02c_rs_minb_*.txt; cat gpurun_out/r02c_quick_minb_*.json | cut -c1-120
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; echo "exit $?" >> gpurun_out/r02c_gputests.log); tail -5 gpurun_out/r02c_gputests.log
