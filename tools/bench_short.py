"""Prints the headline numbers and the top kernels of a bench.py JSON line (tools/: scratch helper for GPU sessions)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"))
for k, v in list(d["kernel_profile"].items())[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("  ", k, v)
