"""Print the headline numbers of a bench.py JSON line (development aid).  usage: python tools/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
      "step_frac", round(d["roofline"]["whole_step_frac"], 3), d["roofline"]["ms_per_launch"])
for k in ("e2e", "e2e_packed", "e2e_inprocess"):
    v = d.get(k)
    if v and "value" in v:
        print(f"{k:14s}", round(v["value"]), "GCUPS", round(v["ms_per_step"], 1), "ms |", v.get("limiter"))
print({k: (round(v["resident_gcups"]), round(v["roofline_frac"], 3)) for k, v in d["modes"].items() if isinstance(v, dict)})
c = d.get("configs") or {}
if "C1" in c:
    print("C1", {k: c["C1"].get(k) for k in ("e2e_gcups", "resident_gcups", "roofline_frac", "oracle_mismatches")}, c["C1"].get("reference_kernels"))
if "C3" in c:
    print("C3", {k: c["C3"].get(k) for k in ("resident_gcups", "roofline_frac_per_gpu", "e2e_gcups", "e2e_ms", "oracle_sample", "e2e_limiter")})
if "C4" in c:
    print("C4", {k: c["C4"].get(k) for k in ("score_resident_gcups", "score_roofline_frac_per_gpu", "score_e2e_gcups", "score_oracle_sample")}, c["C4"].get("align_subset"))
print("cpu_baseline", d.get("cpu_baseline"), "clocks", d.get("clocks"))
