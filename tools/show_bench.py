"""Print the essentials of a bench.py JSON line (argv[1] = log file)."""
import json
import sys

r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c = r["config"]
print("ms/step", round(r["ms_per_step"], 2), "value", f"{r['value']:.3e}", "frac", round(r["roofline"]["frac"], 3),
      "launches", r.get("gpu_launches"))
print("per pass", c.get("per_pass_ms"), "| init:", c.get("init"))
if r.get("zero_support_skipping"):
    print("zero-support", round(r["zero_support_skipping"]["ms_per_step"], 2))
if r.get("e2e"):
    print("e2e", round(r["e2e"]["ms_per_step"], 1), r["e2e"].get("phases_ms"))
