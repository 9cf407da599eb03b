#!/usr/bin/env python3
"""Print the tables of profiles/README.md and DESIGN.md section 8 from a bench.py JSON line.

    python profiles/summarise.py profiles/r01_bench_v12.json [more.json ...]
"""
import json
import sys


def load(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def main():
    for path in sys.argv[1:]:
        d = load(path)
        if d.get("impl") == "reference":
            print(f"{path}: reference arm {d['value']:.1f} {d['unit']}  ({d['cpu_baseline']['sample'][:90]}...)")
            continue
        e2e = d.get("e2e", {})
        print(f"{path}: N={d['n_gpus']}  value {d['value'] / 1e3:.2f} GS/s  {d['ms_per_step']:.3f} ms/step  e2e {e2e.get('value', 0) / 1e3:.2f} GS/s"
              f"  launches/step {d.get('gpu_launches', 0) / max(1, d.get('steps', 1)):.0f}")
        r = d.get("roofline")
        if not r:
            continue
        print(f"  roofline: {r['kernel']}: {r['achieved']:.0f} of {r['peak']:.0f} {r['unit']} = {100 * r['frac']:.1f} %"
              f"  (traffic {r.get('traffic')} B/launch, algorithmic {r.get('algorithmic_bytes_per_launch', 0):.0f} B/launch)")
        ws = r.get("whole_step")
        if ws:
            print(f"  whole batch vs SURVEY 8d bytes: {ws['achieved_gbs']:.0f} GB/s = {100 * ws['frac']:.1f} % ({ws['ms_per_batch']:.4f} ms per batch)")
        # round 1 files say ms_per_step / launches_per_step (a step was one batch of 64 frames), round 2 files ms_per_batch
        ms = lambda v: v.get("ms_per_batch", v.get("ms_per_step"))
        ln = lambda v: v.get("launches_per_batch", v.get("launches_per_step"))
        print(f"  {'kernel':22s} {'ms/batch':>8s} {'launches':>8s} {'alg GB/s':>9s} {'of peak':>8s}")
        for k, v in sorted(r.get("per_kernel", {}).items(), key=lambda kv: -ms(kv[1])):
            gbs = f"{v['achieved_gbs']:.0f}" if "achieved_gbs" in v else "-"
            frac = f"{100 * v['frac']:.0f} %" if "frac" in v else "-"
            print(f"  {k:22s} {ms(v):8.4f} {ln(v):8.1f} {gbs:>9s} {frac:>8s}")
        if "pinned_process" in e2e:
            pp = e2e["pinned_process"]
            print(f"  e2e beside the headline: pinned process() {pp['value'] / 1e3:.2f} GS/s = {100 * pp.get('of_link_bound', 0):.0f} % of the measured link bound")
        if "autocorr_sweep" in d:
            for size, v in d["autocorr_sweep"].items():
                print(f"  autocorr {size}: {v['us_per_autocorrelation']:.2f} us  {v['gbs_vs_bytes_moved']:.0f} GB/s moved ({100 * v['frac_vs_bytes_moved']:.0f} %)  {v['gbs_vs_28N']:.0f} GB/s vs 28N ({100 * v['frac_vs_28N']:.0f} %)")
        pt = e2e.get("plugin_thread_per_block_us")
        if pt:
            print(f"  plugin thread per 2 MiB block: {pt['in_the_plugin_between_callbacks']:.0f} us in the plugin, {pt['inside_the_library_callback']:.0f} us in the library's callback")
        for name, v in (d.get("other_shapes") or {}).items():
            if "value" in v:
                print(f"  other shape {name}: {v['value'] / 1e3:.2f} GS/s  {v['ms_per_batch']:.4f} ms per batch  frame {v.get('frame')}")
        sb = d.get("superbandwidth")
        if sb and "ms_per_stitch" in sb:
            fr = sb.get("frames") or {}
            par = sb.get("parity_vs_one_gpu_path") or {}
            print(f"  superbandwidth, {sb['hops']} hops of {sb['n_per_hop']}: {sb['ms_per_stitch']:.4f} ms per stitch, one GPU {sb.get('one_gpu_ms', float('nan')):.4f} ms "
                  f"-> x{sb.get('speedup_vs_one_gpu', float('nan')):.2f}; stitch + frames {fr.get('ms_per_round_stitch_plus_frames')} ms; "
                  f"max err / peak {par.get('max_err_over_peak')}, lags equal {par.get('lags_equal')}; NCCL all-gather formulation {sb.get('nccl_allgather_baseline_ms')} ms")
        elif sb:
            print(f"  superbandwidth: {json.dumps(sb)[:400]}")
        for key in ("e2e_int8_transport", "variants"):
            if key in d:
                print(f"  {key}: {json.dumps(d[key])[:400]}")


if __name__ == "__main__":
    main()
