#!/usr/bin/env python
"""Print the key fields of a bench.py JSON line read from stdin:  python bench.py ... | python tools/benchline.py [label]"""
import json
import sys

d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
print(" ".join(sys.argv[1:]), f"value={d['value']:.4g} {d['unit']}", f"ms/step={d['ms_per_step']:.1f}",
      f"sweep_GBps={r.get('achieved') or 0:.0f}", f"sweep_ms={(d['ms_per_step'] * (r.get('sweep_share_of_step') or 0)):.1f}",
      f"e2e={e.get('value') or 0:.4g}", f"states/step={d.get('states_swept_per_step')}")
