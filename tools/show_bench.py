"""Condensed view of a bench.py JSON line (python tools/show_bench.py FILE)."""
import json, sys
d = json.loads([ln for ln in open(sys.argv[1]).read().splitlines() if ln.startswith("{")][-1])
for k in ("value", "ms_per_step", "launch", "gpu_launches"):
    print(k, d.get(k))
e = d["e2e"]
print("e2e", e["value"], e["ms_per_step"], "serial", e["serial"]["value"], "idx_only", e.get("idx_only", {}).get("value"), "numa", e.get("numa"))
print("kernels_ms", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["kernels_ms"].items() if k != "GBps"})
print("inflight", d.get("device_batches_in_flight"))
print("refcuda", d.get("reference_cuda"))
print("cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "sample"})
print("clocks", d.get("clocks"))
c = d.get("configs") or {}
print(c.get("what"), c.get("error"))
for r in c.get("rows", []):
    if len(sys.argv) > 2 and sys.argv[2] not in r["config"]:
        continue
    extra = {k: round(v, 3) for k, v in r.items() if k in ("us_per_iter", "fp32_issue_frac_of_gpu", "points_per_s_all_gpus")}
    print(f"{r['config']:26s} {r['kernel'][:66]:66s} {r['ms']:9.4f} ms {r['GBps_per_gpu']:8.1f} GB/s {r['frac_of_peak']:.3f}", extra)
