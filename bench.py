#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE configs[1] -- 1080p60 geometry (1125 total lines, the GUI's convention;
W = 740), 25 MS/s float32 IQ.  One step = one batch of 64 frames' worth of synthetic IQ (640 decimator blocks of
41666 pairs = 26.7 M pairs = 213 MB, larger than L2) through the whole hot path:
    fused demod+resample -> pixel stream -> frame stage (auto-gain, temporal IIR, collapse, sync search, re-centre)
    and, beside it, the frame-rate detector: every capture of 3.1*fs/55 samples is demodulated, autocorrelated
    (2^20-point FFT + IFFT) and accumulated into the two lag plots.
`value`   MS/s with the IQ already resident in HBM (CUDA events around exactly K steps, max over ranks).
`e2e`     the same metric through the reference-facing call: tsdrgpu_pipeline_process() == the plugin's process()
          callback with HOST buffers (pinned), H2D of every block and D2H of every finished frame inside the timed
          region (host wall clock bracketed by device synchronisation, max over ranks).
N > 1     N independent streams, one per GPU (the path has no cross-stream exchange: replicas, weak scaling).
--impl reference   the reference's own threaded CPU pipeline (oracle/_ref: libTSDRLibrary.so + its RawFile plugin
          with pacing off) on this box's host cores, same geometry; falls back to the pinned C port when the
          reference binary is absent.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, HEIGHT, FV = 25_000_000, 1125, 60.0
FRAMES_PER_STEP = 64
METRIC = "IQ MS/s ingested -> 1080p60 frames (demod+resample+frame stage+autocorrelation), whole job"


def geometry():
    from tempestsdr_b200 import _native as N
    w, pr, pt = C.c_int(0), C.c_double(0), C.c_double(0)
    N.lib().tsdrgpu_geometry(FS, HEIGHT, FV, C.byref(w), C.byref(pr), C.byref(pt))
    return w.value


def make_iq(pairs: int, seed: int) -> np.ndarray:
    """Video-like synthetic IQ, generated for one frame period and tiled (cheap, deterministic)."""
    from tempestsdr_b200 import synth
    per_frame = int(FS / FV)
    base = synth.video_like_iq(4 * per_frame, FS, 2576, 1125, FV, seed=seed, snr_db=25.0)
    reps = (2 * pairs + base.size - 1) // base.size
    return np.ascontiguousarray(np.tile(base, reps)[: 2 * pairs])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        """index: one GPU index, or a comma-separated list (then `per_gpu` in the result tells the GPUs apart)."""
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, per = [], [], set(), {}
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                sm.append(float(p[1])); mx.append(float(p[2]))
                per.setdefault(p[0], []).append(float(p[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
            if len(per) > 1:
                out["per_gpu"] = {k: statistics.median(v) for k, v in sorted(per.items())}
        return out


# --------------------------------------------------------------------------------------------------- our arm
class DeviceStep:
    """One step of the hot path with the IQ resident in HBM."""

    def __init__(self, gpu, iq_dev, w):
        import torch
        from tempestsdr_b200.api import FrameRateDetector, PostProcessFlags
        self.torch, self.gpu, self.iq, self.w = torch, gpu, iq_dev, w
        self.block = int(0.1 * FS / FV)
        self.nblocks = FRAMES_PER_STEP * 10
        self.n = w * HEIGHT
        self.up = w * HEIGHT * FV
        self.rs = gpu.resampler()
        self.pp = gpu.post_processor()
        self.pp.set_overlap(not os.environ.get("BENCH_NO_OVERLAP"))   # sync search + re-centring of batch k under the kernels of batch k+1
        self.frd = gpu.framerate_detector()
        self.flags = PostProcessFlags(autoshift=True, lowpass_before_sync=True)     # the GUI's defaults
        self.cap = FrameRateDetector.capture_size(FS)
        max_pix = int(self.rs.plan((self.block, self.nblocks), self.up, float(FS))) + 1024
        self.pix = torch.empty(max_pix + self.n + 1024, dtype=torch.float32, device=iq_dev.device)
        self.pix_fill = 0
        self.frames_out = [torch.empty(FRAMES_PER_STEP * self.n, dtype=torch.float32, device=iq_dev.device) for _ in range(2)]
        self.pairs = self.block * self.nblocks
        self.mag = torch.empty(self.cap + self.pairs, dtype=torch.float32, device=iq_dev.device)   # demodulated stream, capture-aligned
        self.mag_fill = 0
        self.frames = 0
        self.captures = 0
        self.k = 0
        self.chunk = int(os.environ.get("BENCH_FRAME_CHUNK", "0"))

    def __call__(self):
        gpu = self.gpu
        # samples -> pixels (fused demod + resample), appended behind the pixels left over from the last step
        # (the same pass also leaves the magnitudes for the frame-rate detector: one demodulation feeds both consumers,
        # as am_demod does in the reference's process(), TSDRLibrary.c:286-292)
        fused_mag = not os.environ.get("BENCH_SEPARATE_DEMOD")
        out = self.rs.process(self.iq, (self.block, self.nblocks), self.up, float(FS), in_is_iq=True, out=self.pix[self.pix_fill:],
                              mag_out=self.mag[self.mag_fill:] if fused_mag else None)
        self.pix_fill += out.numel()
        nf = min(self.pix_fill // self.n, FRAMES_PER_STEP)
        # the frame stage in sub-batches of `chunk` frames (default: the whole step at once).  Smaller sub-batches keep a batch's
        # intermediate frames in L2 between the kernels of the stage at the price of more launches.
        chunk = self.chunk if self.chunk > 0 else nf
        fo = self.frames_out[self.k & 1]
        for c0 in range(0, nf, chunk):
            c1 = min(nf, c0 + chunk)
            self.pp.process(self.pix[c0 * self.n: c1 * self.n], self.w, HEIGHT, 0.0, 0.1, self.flags, out=fo[c0 * self.n: c1 * self.n], want_results=False)
        self.k += 1
        left = self.pix_fill - nf * self.n
        if left:
            self.pix[:left].copy_(self.pix[nf * self.n: self.pix_fill])      # left << nf*n: the ranges do not overlap
        self.pix_fill = left
        self.frames += nf
        # frame-rate detector: the whole stream is demodulated once; every complete capture of 3.1*fs/55 samples is
        # autocorrelated (batched FFTs) and accumulated in order
        if not fused_mag:
            gpu.chk(gpu._lib.tsdrgpu_am_demod(gpu._h, gpu.stream, self.iq.data_ptr(), self.pairs, self.mag.data_ptr() + 4 * self.mag_fill))
        self.mag_fill += self.pairs
        ncap = self.mag_fill // self.cap
        if ncap:
            self.frd.run_batch(FS, self.mag, self.cap, ncap, self.cap)
            rest = self.mag_fill - ncap * self.cap
            if rest:
                self.mag[:rest].copy_(self.mag[ncap * self.cap: self.mag_fill])
            self.mag_fill = rest
            self.captures += ncap

    def join(self):
        self.pp.join()


def collect_profile(gpu):
    names = C.create_string_buffer(48 * 48)
    tot = (C.c_double * 48)(); cnt = (C.c_uint64 * 48)(); n = C.c_int(0)
    gpu.chk(gpu._lib.tsdrgpu_profile_collect(gpu._h, names, tot, cnt, 48, C.byref(n)))
    out = {}
    for i in range(n.value):
        nm = names.raw[48 * i: 48 * (i + 1)].split(b"\0")[0].decode()
        out[nm] = (tot[i], cnt[i])
    return out


def cpu_baseline(w, seconds_budget=25.0):
    """The same work as one GPU step on ONE host core: the compiled reference's stage functions driven serially
    (oracle/_ref when present, else the pinned C port).  Bounded sample: 8 frames of IQ + 2 autocorrelation captures."""
    from oracle import oracle as orc
    O = orc.best()
    block = int(0.1 * FS / FV)
    nframes = 8
    pairs = block * 10 * nframes
    iq = make_iq(pairs, seed=77)
    cap = int(3.1 * FS / 55.0)
    t0 = time.perf_counter()
    mag = O.am_demod(iq)
    rs = O.resampler()
    pix = np.concatenate([rs.run(mag[k * block:(k + 1) * block], w * HEIGHT * FV, float(FS)) for k in range(10 * nframes)])
    pp = O.postprocessor(FS, HEIGHT, FV, 1, 0)
    n = w * HEIGHT
    for k in range(pix.size // n):
        pp.run(pix[k * n:(k + 1) * n], w, HEIGHT, 0.0, 0.1, 1, 0)
    t_stream = time.perf_counter() - t0
    det = O.framerate_detector()
    ncap = 2
    capdata = np.tile(mag, (ncap * cap + mag.size - 1) // mag.size)[: ncap * cap]
    t1 = time.perf_counter()
    for c in range(ncap):
        det.run(FS, capdata[c * cap:(c + 1) * cap])
    t_cap = (time.perf_counter() - t1) / ncap
    # one GPU step does pairs_step samples of the stream path and pairs_step/cap captures
    per_sample = t_stream / pairs + t_cap / cap
    return {"value": 1e-6 / per_sample, "unit": "MS/s", "cores": 1, "kind": "reference" if O.kind == "reference" else "port",
            "sample": f"{nframes} frames ({pairs} IQ pairs) through demod+resample+frame stage: {t_stream:.2f} s; "
                      f"{ncap} captures of {cap} samples autocorrelated: {t_cap:.2f} s each; stages driven serially on one thread"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tempestsdr_b200 import api, pipeline

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"            # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gpu = api.Context(local)
    w = geometry()
    block = int(0.1 * FS / FV)
    pairs = block * 10 * FRAMES_PER_STEP
    iq_host = make_iq(pairs, seed=1000 + rank)
    iq_pinned = torch.from_numpy(iq_host).pin_memory()
    iq_dev = iq_pinned.cuda(non_blocking=True)
    step = DeviceStep(gpu, iq_dev, w)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = gpu.launches
    frames0, caps0 = step.frames, step.captures
    sampler = ClockSampler(local)
    sampler_all = ClockSampler(",".join(str(i) for i in range(world))) if world > 1 else None     # every GPU of the job, informational
    if rank == 0:
        sampler.start()
        if sampler_all:
            sampler_all.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps     # host time to ENQUEUE a step (nothing is waited for)
    step.join()                                   # the side stream's tail (last sync search + re-centring) is inside the timed region
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and sampler_all:
        allg = sampler_all.stop()
        if allg.get("sm_mhz") is not None:
            clocks["all_gpus"] = allg                 # the headline keys above stay rank 0's GPU, as at N = 1
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    launches = gpu.launches - launches0
    frames_done, caps_done = step.frames - frames0, step.captures - caps0

    if os.environ.get("BENCH_QUICK"):                 # used under ncu and for the opt-in variants: the timed steps only
        quick = {"quick": True, "value": world * args.steps * pairs / (ms_total * 1e-3) / 1e6, "unit": "MS/s", "ms_per_step": ms_total / args.steps}
        if os.environ.get("BENCH_VARIANT_CHECK") and os.environ.get("TSDRGPU_AUTOCORR_HALF") and world == 1:
            # the batched frame-rate detector with and without the half-size transforms on the same four captures: how far the two
            # running-mean plots are apart, relative to the plot's peak (the parity bound of the default path is 1e-5)
            try:
                step.join(); torch.cuda.synchronize()
                caps = torch.abs(torch.randn(4 * step.cap, device=iq_dev.device)) + 0.25
                got = {}
                for mode in ("1", None):
                    if mode is None:
                        os.environ.pop("TSDRGPU_AUTOCORR_HALF", None)
                    det = gpu.framerate_detector()
                    det.run_batch(FS, caps, step.cap, 4, step.cap)
                    (_, fp), (_, lp) = det.plots(FS)
                    got[mode] = (fp.copy(), lp.copy())
                os.environ["TSDRGPU_AUTOCORR_HALF"] = "1"
                quick["frd_plot_max_rel_diff_vs_default"] = max(float(np.max(np.abs(got["1"][i] - got[None][i])) / np.max(np.abs(got[None][i]))) for i in (0, 1))
            except Exception as e:
                quick["frd_plot_check_failed"] = repr(e)[:160]
        if rank == 0:
            emit(quick)
        if world > 1:
            dist.destroy_process_group()
        return
    # host cost of enqueueing a step when nothing throttles it (3 steps right after a full synchronisation: the resampler's
    # 4-slot descriptor ring cannot be full yet).  Informational: it tells how far the host is from being the bottleneck.
    step.join(); torch.cuda.synchronize()
    t_h = time.perf_counter()
    for _ in range(3):
        step()
    host_enqueue_free_ms = (time.perf_counter() - t_h) * 1e3 / 3
    step.join(); torch.cuda.synchronize()
    # ---- per-kernel timing pass (separate from the timed region above): CUDA events on the launching stream
    # The side stream is switched off for this pass so that every kernel's event pair measures that kernel alone
    # (with it on, intervals on the main stream also contain the slowdown from sharing the chip with the sync search).
    step.join(); torch.cuda.synchronize()
    step.pp.set_overlap(False)
    gpu.chk(gpu._lib.tsdrgpu_profile_enable(gpu._h, 1))
    collect_profile(gpu)
    prof_steps = 3
    caps_before_prof = step.captures
    for _ in range(prof_steps):
        step()
    prof = collect_profile(gpu)
    gpu.chk(gpu._lib.tsdrgpu_profile_enable(gpu._h, 0))
    step.pp.set_overlap(not os.environ.get("BENCH_NO_OVERLAP"))

    # ---- e2e through the C-ABI pipeline with host buffers
    chunk = 512 * 1024 * 8                     # floats per process() call (8x the RawFile plugin's block)
    pl = pipeline.Pipeline(samplerate=FS, height=HEIGHT, refreshrate=FV, batch_frames=16, batch_blocks=160, block_when_busy=True,
                           device=local, params={"autoshift": 1, "lowpass_before_sync": 1})
    host = iq_pinned.numpy()
    base_ptr = iq_pinned.data_ptr()

    def feed_once():
        pos = 0
        while pos < host.size:
            n = min(chunk, host.size - pos)
            pl.process_ptr(base_ptr + 4 * pos, n, 0)
            pos += n

    # the link under the e2e number: pinned H2D and D2H of one step's bytes, alone and together (context, not a claim)
    def link_gbs():
        d_in = torch.empty_like(iq_dev); h_out = torch.empty(FRAMES_PER_STEP * step.n, dtype=torch.float32).pin_memory()
        d_out = step.frames_out[0]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        res = {}
        reps = 4
        for name, both in (("h2d", (True, False)), ("d2h", (False, True)), ("duplex", (True, True))):
            for timed in (False, True):                      # one untimed pass first (first touch of the pinned pages, stream creation)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(reps if timed else 1):
                    if both[0]:
                        with torch.cuda.stream(s1):
                            d_in.copy_(iq_pinned, non_blocking=True)
                    if both[1]:
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
            nbytes = (iq_pinned.numel() * 4 if both[0] else 0) + (h_out.numel() * 4 if both[1] else 0)
            res[name + "_gbs"] = nbytes / dt / 1e9
        return res
    link = link_gbs()
    e2e_steps = max(2, min(args.steps, 6))
    feed_once(); pl.flush()
    barrier()
    s0 = pl.stats()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        feed_once()
    pl.flush()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    s1 = pl.stats()
    e2e_val = world * e2e_steps * pairs / t_e2e.item() / 1e6
    h2d = (s1.h2d_bytes - s0.h2d_bytes) // e2e_steps
    d2h = (s1.d2h_bytes - s0.d2h_bytes) // e2e_steps
    e2e_frames = s1.frames_delivered - s0.frames_delivered
    pl.close()

    # ---- the same e2e with the samples crossing PCIe as int8 (raw sink / tsdrgpu_pipeline_process_raw):
    # reported beside the headline, never instead of it (the reference arm reads float32)
    e2e_int8 = None
    if world == 1 and not os.environ.get("BENCH_NO_INT8"):
        q8 = torch.clamp(torch.round(iq_pinned * (100.0 / float(iq_pinned.abs().max()))), -127, 127).to(torch.int8).pin_memory()
        pl8 = pipeline.Pipeline(samplerate=FS, height=HEIGHT, refreshrate=FV, batch_frames=16, batch_blocks=160, block_when_busy=True,
                                device=local, params={"autoshift": 1, "lowpass_before_sync": 1})
        ptr8, n8 = q8.data_ptr(), q8.numel()

        def feed8():
            pos = 0
            while pos < n8:
                n = min(chunk, n8 - pos)
                pl8.process_raw_ptr(ptr8 + pos, 1, n, 0)
                pos += n
        feed8(); pl8.flush()
        a0 = pl8.stats(); t8 = time.perf_counter()
        for _ in range(e2e_steps):
            feed8()
        pl8.flush(); torch.cuda.synchronize()
        t8 = time.perf_counter() - t8
        a1 = pl8.stats()
        e2e_int8 = {"value": e2e_steps * pairs / t8 / 1e6, "unit": "MS/s", "h2d_bytes_per_step": int((a1.h2d_bytes - a0.h2d_bytes) // e2e_steps),
                    "d2h_bytes_per_step": int((a1.d2h_bytes - a0.d2h_bytes) // e2e_steps), "frames_delivered": int(a1.frames_delivered - a0.frames_delivered),
                    "how": "tsdrgpu_pipeline_process_raw(int8) on pinned host samples, converted on the device (TSDRPlugin_RawFile.c:247 values); "
                           "float32 frames copied back as in e2e"}
        pl8.close()

    # ---- N > 1 only: the path's one real exchange, the superbandwidth stitch with one hop per GPU (configs[3])
    superb = None
    if world > 1:
        from tempestsdr_b200 import superband
        hop_pairs = 10 * int(FS / FV)                        # SUPER_SAMPLES_TO_RECORD frames per hop -> N = 2^22
        hop = iq_dev[: 2 * hop_pairs].contiguous()
        for _ in range(2):
            superband.stitch_distributed(gpu, hop, int(FS / FV))
        barrier()
        s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        s0e.record()
        for _ in range(reps):
            res, lags, n_fft = superband.stitch_distributed(gpu, hop, int(FS / FV))
        s1e.record()
        barrier()
        tms = torch.tensor([s0e.elapsed_time(s1e) / reps], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        superb = {"hops": world, "n_per_hop": n_fft, "ms_per_stitch": tms.item(), "stitched_MS_per_s": world * n_fft / (tms.item() * 1e-3) / 1e6,
                  "allgather_bytes_per_rank": 8 * (n_fft + n_fft // 2), "collective": "one NCCL all_gather_into_tensor per stitch"}
        # the same stitch with the all-gather fused into the forward transforms' last pass (NVLink peer stores through CUDA IPC,
        # a 4-byte all-reduce as the barrier): DESIGN.md section 6
        try:
            ex = superband.PeerExchange(gpu, 2 * n_fft)
            for _ in range(2):
                superband.stitch_distributed_fused(gpu, hop, int(FS / FV), ex)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                res_f, lags_f, _ = superband.stitch_distributed_fused(gpu, hop, int(FS / FV), ex)
            f1.record()
            barrier()
            tf = torch.tensor([f0.elapsed_time(f1) / reps], device="cuda")
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            same = torch.tensor([int(bool(torch.equal(res, res_f)) and list(lags_f) == list(lags))], device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            superb["fused_peer_store"] = {"ms_per_stitch": tf.item(), "stitched_MS_per_s": world * n_fft / (tf.item() * 1e-3) / 1e6,
                                          "bit_identical_to_nccl_path": bool(same.item()),
                                          "how": "FFT last pass stores into every rank's gather buffer (peer memory), then a 4-byte all-reduce as barrier"}
            ex.close()
        except Exception as e:                                   # peer access unavailable on this box: the NCCL figure stands alone
            superb["fused_peer_store"] = {"unavailable": repr(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (by total device time in the profiled steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    ratio = w * HEIGHT * FV / FS
    # algorithmic bytes per launch (DESIGN.md "algorithmic bytes"), per kernel
    n_pix = step.n * FRAMES_PER_STEP
    prof_caps = (step.captures - caps_before_prof) / prof_steps          # captures autocorrelated per profiled step
    alg = {   # ALGORITHMIC bytes per launch
        "rs_main": pairs * (8 + 4 * ratio + (0 if os.environ.get("BENCH_SEPARATE_DEMOD") else 4)),   # 8 B per IQ pair in + 4 B per pixel out (+ 4 B magnitude out)
        "fs_minmax": 4 * n_pix, "fs_normalise": 8 * n_pix, "fs_timelowpass": 8 * n_pix, "fs_norm_lowpass": 8 * n_pix,
        "fs_collapse": 4 * n_pix, "fs_shift": 8 * n_pix, "demod_kernel": 12 * pairs,
        # one FFT pass reads and writes every complex point once; a launch covers all captures of the step (grid.y);
        # a step's autocorrelations are 4 launches (2 passes forward, 2 inverse)
        "fft_pass_kernel": 16 * (1 << 20) * prof_caps * 4 / max(1.0, prof.get("fft_pass_kernel", (0, 4 * prof_steps))[1] / prof_steps),
    }
    label = {"rs_main": "rs_main<IQ> (fused demod+resample)", "fft_pass_kernel": "fft_pass_kernel (one pass over every capture of the step)"}
    total_prof = sum(t for t, _ in prof.values()) or 1.0
    kernels = {k: {"ms_per_step": t / prof_steps, "launches_per_step": c / prof_steps, "share": t / total_prof} for k, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    dom = next(iter(kernels))
    # the roofline object is for the dominant kernel of the step among the bandwidth kernels; fs_sync (one cluster of 8
    # CTAs, FP64-latency bound by construction) is listed in per_kernel but has no bandwidth roofline
    roof_k = next((k for k in kernels if alg.get(k)), "rs_main")
    t_k, c_k = prof.get(roof_k, (0.0, 0))
    achieved = alg[roof_k] / (t_k / c_k * 1e-3) / 1e9 if c_k else None
    roofline = {"kernel": label.get(roof_k, roof_k), "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[roof_k], "avg_launch_ms": (t_k / c_k) if c_k else None,
                "dominant_kernel_by_time": dom,
                "per_kernel": {k: dict(v, **({"achieved_gbs": alg[k] * v["launches_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9,
                                              "frac": alg[k] * v["launches_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9 / peak} if alg.get(k) else {})) for k, v in kernels.items()}}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
        roofline["traffic"] = tr.get(roof_k, {}).get("dram_bytes_per_launch")
        roofline["traffic_source"] = tr.get(roof_k, {}).get("source")
    except Exception:
        pass
    value = world * args.steps * pairs / (ms_total * 1e-3) / 1e6
    # opt-in code paths measured on the same workload in a child process (own CUDA context: whatever happens there cannot touch
    # the numbers above).  Informational; the headline is always the default path.
    variants = None
    if world == 1 and not os.environ.get("BENCH_NO_VARIANTS"):
        variants = {}
        for name, env in (("autocorr_half_size", {"TSDRGPU_AUTOCORR_HALF": "1"}),):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(args.steps), "--warmup", str(max(args.warmup, 3))],
                                   capture_output=True, text=True, timeout=180, env=dict(os.environ, BENCH_QUICK="1", BENCH_VARIANT_CHECK="1", **env))
                q = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
                variants[name] = {"value": q["value"], "unit": "MS/s", "ms_per_step": q["ms_per_step"], "env": env,
                                  **{k: q[k] for k in ("frd_plot_max_rel_diff_vs_default", "frd_plot_check_failed") if k in q},
                                  "note": "device-resident value of the same workload with this opt-in path (DESIGN.md section 9)"}
            except Exception as e:
                variants[name] = {"unavailable": repr(e)[:160], "env": env}
    cpu = cpu_baseline(w)
    line = {
        "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (f64 accumulators where the reference uses them)", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 1080p60 geometry (1125 total lines -> 740x1125 px frames), 25 MS/s float32 IQ, "
                               f"{FRAMES_PER_STEP} frames per step ({pairs} IQ pairs, {8 * pairs / 1e6:.0f} MB > L2, so no L2 flush is needed)",
                   "frames_per_step": frames_done / args.steps, "autocorr_captures_per_step": caps_done / args.steps,
                   "frames_per_s": world * frames_done / (ms_total * 1e-3), "parallelism": f"replicas x{world}" if world > 1 else "single stream",
                   "host_enqueue_ms_per_step": host_enqueue_ms, "host_enqueue_ms_per_step_unthrottled": host_enqueue_free_ms,
                   "flags": "AUTOSHIFT=1, LOW_PASS_BEFORE_SYNC=1, AUTOGAIN_AFTER=0, motionblur 0 (GUI defaults), PLL write-back off"},
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": "MS/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "frames_delivered": int(e2e_frames), "pcie_link_measured": link,
                # bytes per sample over the link: 8 in + 4*pixels-per-sample out; bound by each direction alone and by both together
                "link_bound_MS_per_s": min(link["h2d_gbs"] / 8.0, link["d2h_gbs"] / (4.0 * step.n * FRAMES_PER_STEP / pairs),
                                           link["duplex_gbs"] / (8.0 + 4.0 * step.n * FRAMES_PER_STEP / pairs)) * 1e3,
                "how": "tsdrgpu_pipeline_process() on pinned host IQ in 16 MiB calls, "
                "frames copied back to pinned host slots; host wall clock between device synchronisations"},
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
    }
    if superb:
        line["superbandwidth"] = superb
    if e2e_int8:
        line["e2e_int8_transport"] = e2e_int8
    if variants:
        line["variants"] = variants
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    w = geometry()
    ncores = os.cpu_count() or 1
    if not orc.have_ref():
        cpu = cpu_baseline(w)
        cpu["sample"] = "reference binary absent: pinned C port, " + cpu["sample"]
        emit(({"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "higher_is_better": True, "data": "synthetic", "cpu_baseline": cpu,
                          "config": {"workload": "BASELINE configs[1] geometry, C port of the reference stages, one thread"},
                          "e2e": {"value": cpu["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    # The reference's own threaded pipeline is measured in a CHILD process: its worker threads race on shared state (SURVEY F9;
    # TSDRLibrary.c:62-94 leaves fields uninitialised) and now and then the unmodified library segfaults during start-up.
    # A crash must not cost the round its reference number: retry, and only then fall back to the stage-driven figure.
    child = None
    per_frame = int(FS / FV)
    tmp = tempfile.NamedTemporaryFile(prefix="tsdr_iq_", suffix=".raw", delete=False)     # one recording for every attempt
    make_iq(16 * per_frame, seed=1000).tofile(tmp); tmp.close()
    tries = 10
    try:
        for attempt in range(tries):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference_child", "--gpus", str(args.gpus),
                                "--steps", str(args.steps), "--warmup", str(args.warmup)], capture_output=True, text=True,
                               env=dict(os.environ, BENCH_REF_IQ_FILE=tmp.name))
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and lines:
                child = json.loads(lines[-1]); child["attempts"] = attempt + 1
                break
            sys.stderr.write(f"[bench] reference pipeline attempt {attempt + 1} ended with rc={r.returncode} (the unmodified library crashed); retrying\n")
            time.sleep(0.2 * (attempt + 1))
    finally:
        os.unlink(tmp.name)
    if child is None:
        cpu = cpu_baseline(w)
        cpu["sample"] = f"the reference's threaded pipeline crashed {tries} times in a row; its stage functions driven serially instead: " + cpu["sample"]
        emit(({"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "cpu_baseline": cpu, "config": {"workload": "BASELINE configs[1] geometry, the compiled reference's stage functions, one thread"},
               "e2e": {"value": cpu["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    fps, caps, seconds, per_frame = child["fps"], child["caps"], child["seconds"], child["per_frame"]
    value = fps * per_frame / 1e6
    cpu = {"value": value, "unit": "MS/s", "cores": min(ncores, 6), "kind": "reference",
           "sample": f"the reference's own threaded pipeline (plugin + decimate + post-process + video + autocorr threads) for "
                     f"{args.steps} x {seconds:.0f} s on {ncores} host cores; counts FRAMES DELIVERED x samples per frame "
                     f"(it drops whole blocks when a ring is full); {fps:.1f} frames/s, {caps:.2f} autocorrelation captures/s"
                     + (f"; attempt {child['attempts']} (earlier ones crashed inside the reference library)" if child["attempts"] > 1 else "")}
    emit(({"impl": "reference", "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": seconds * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "cpu_baseline": cpu,
                      "config": {"workload": "BASELINE configs[1]: 1080p60 geometry (1125 lines), 25 MS/s float32 IQ from a file through "
                                             "TSDRPlugin_RawFile (pacing off) and the unmodified reference library", "frames_per_s": fps},
                      "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_reference_child(args):
    """One attempt at driving the unmodified reference library (see run_reference); prints {"fps", "caps", ...} as JSON."""
    from oracle import oracle as orc
    lib = C.CDLL(orc.REF_LIB_SO)
    per_frame = int(FS / FV)
    iq_file = os.environ["BENCH_REF_IQ_FILE"]              # written (and removed) by the parent
    FRAME_CB = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_void_p)
    VALUE_CB = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_double, C.c_void_p)
    PLOT_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_uint32, C.c_void_p)
    count = {"frames": 0, "plots": 0}
    fcb = FRAME_CB(lambda b, ww, hh, c: count.__setitem__("frames", count["frames"] + 1))
    vcb = VALUE_CB(lambda i, a, b, c: None)
    pcb = PLOT_CB(lambda p, o, v, s, sr, c: count.__setitem__("plots", count["plots"] + 1))
    t = C.c_void_p()
    lib.tsdr_init(C.byref(t), vcb, pcb, None)
    lib.tsdr_setresolution.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.tsdr_motionblur.argtypes = [C.c_void_p, C.c_float]
    lib.tsdr_setgain.argtypes = [C.c_void_p, C.c_float]
    lib.tsdr_setparameter_int.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    lib.tsdr_loadplugin.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    lib.tsdr_readasync.argtypes = [C.c_void_p, FRAME_CB, C.c_void_p]
    lib.tsdr_stop.argtypes = [C.c_void_p]
    lib.tsdr_setresolution(t, HEIGHT, FV); lib.tsdr_motionblur(t, 0.0); lib.tsdr_setgain(t, 0.5)
    for pid, v in ((0, 1), (1, 0), (6, 1)):          # AUTOSHIFT=1, PLL=0, LOW_PASS_BEFORE_SYNC=1
        lib.tsdr_setparameter_int(t, pid, v)
    rc = lib.tsdr_loadplugin(t, orc.REF_RAWFILE_NOPACE_SO.encode(), f'"{iq_file}" {FS} float'.encode())
    assert rc == 0, f"tsdr_loadplugin rc={rc}"
    th = threading.Thread(target=lambda: lib.tsdr_readasync(t, fcb, None), daemon=True)
    th.start()
    seconds = 4.0
    time.sleep(1.0)                                  # warm-up: rings grow, first frames arrive
    results = []
    for s in range(args.warmup + args.steps):
        f0, p0, t0 = count["frames"], count["plots"], time.perf_counter()
        time.sleep(seconds)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            results.append(((count["frames"] - f0) / dt, (count["plots"] - p0) / 2 / dt))
    fps = statistics.mean(r[0] for r in results)
    caps = statistics.mean(r[1] for r in results)
    emit({"fps": fps, "caps": caps, "seconds": seconds, "per_frame": per_frame})
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)      # the measurement is over: leave without tsdr_stop / interpreter teardown (the reference's shutdown path races too)


_RESULT_OUT = None


def _claim_stdout():
    """Rank 0 must print exactly ONE line on stdout.  Libraries write there too (NCCL's version banner goes to fd 1 whatever
    NCCL_DEBUG says once it is at least VERSION), so the real stdout is set aside for the result line and fd 1 is pointed at
    stderr for everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_child"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_child":
        run_reference_child(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
