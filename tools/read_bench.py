"""Print the key fields of a bench.py JSON line (bring-up helper)."""
import json
import sys

for path in sys.argv[1:]:
    lines = [x for x in open(path) if x.startswith("{")]
    if not lines:
        print(path, "no JSON line")
        continue
    d = json.loads(lines[-1])
    print(path)
    for k in ("value", "ms_per_step", "algorithmic_tflops", "e2e", "gpu_launches", "clocks", "sustained", "parity"):
        if k in d:
            print("  ", k, d[k])
    r = d.get("roofline", {})
    print("   roofline", {k: r.get(k) for k in ("achieved", "frac", "us_per_launch", "traffic")})
    ex = d.get("extras", {})
    print("   invert", ex.get("invert_ms_all_layers"), " predictive", ex.get("posterior_predictive", {}).get("value"))
    if "cpu_baseline" in d:
        print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
