#!/usr/bin/env python
"""Print the headline fields of a bench.py JSON line: show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d.get("value"), d.get("unit"), "| ms/step", d.get("ms_per_step"), "| e2e", (d.get("e2e") or {}).get("value"),
      "| n_gpus", d.get("n_gpus"))
r = d.get("roofline") or {}
print("roofline frac", r.get("frac"), "achieved", r.get("achieved"), r.get("unit"))
for k, v in (r.get("families") or {}).items():
    print("  ", k, v)
