#!/usr/bin/env python
"""Print a markdown table of the bench lines committed under profiles/ (one row per *.json produced by bench.py)."""
import glob
import json
import os
import sys

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
pat = sys.argv[1] if len(sys.argv) > 1 else "r2*"
print("| file | workload | GPUs | value MB/s | e2e MB/s | CPU ref MB/s (1 thread) | parse / MTF ms | rounds per window | re-parsed blocks |")
print("|---|---|---|---|---|---|---|---|---|")
for f in sorted(glob.glob(os.path.join(root, pat + ".json"))):
    try:
        with open(f) as g:
            d = json.loads(g.read().strip().splitlines()[-1])
    except Exception:
        continue
    if "metric" not in d or d.get("impl") == "reference":
        continue
    pc = d.get("parse_counters") or {}
    km = d.get("kernel_ms") or {}
    print("| `%s` | %s | %d | %.1f | %.1f | %s | %s / %s | %s | %s |" % (
        os.path.basename(f), d["config"]["workload"].replace("|", "/"), d["n_gpus"], d["value"], d["e2e"]["value"],
        (d.get("cpu_baseline") or {}).get("value", "-"), km.get("parse", "-"), km.get("mtf", "-"),
        pc.get("rounds_per_window", "-"), pc.get("reparsed_blocks", "-")))
