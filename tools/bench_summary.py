#!/usr/bin/env python
"""Compact one-line summary of a bench.py JSON line read from stdin (for A/B experiments)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = {k: round(v, 3) for k, v in d.get("stage_ms", {}).items() if v > 0}
print(sys.argv[1] if len(sys.argv) > 1 else "", "ms/step", round(d["ms_per_step"], 4), "tests/s %.4g" % d["value"],
      "e2e_ms", round(d["e2e"]["ms_per_step"], 3), "resolved %.4g" % d["resolved"]["value"], st)
