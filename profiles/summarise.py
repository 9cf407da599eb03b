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
        print(f"  {'kernel':22s} {'ms/step':>8s} {'launches':>8s} {'alg GB/s':>9s} {'of peak':>8s}")
        for k, v in sorted(r.get("per_kernel", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
            gbs = f"{v['achieved_gbs']:.0f}" if "achieved_gbs" in v else "-"
            frac = f"{100 * v['frac']:.0f} %" if "frac" in v else "-"
            print(f"  {k:22s} {v['ms_per_step']:8.4f} {v['launches_per_step']:8.1f} {gbs:>9s} {frac:>8s}")
        for key in ("e2e_int8_transport", "superbandwidth", "variants"):
            if key in d:
                print(f"  {key}: {json.dumps(d[key])[:400]}")


if __name__ == "__main__":
    main()
