#!/usr/bin/env python
"""One line per bench JSON file: value, ms per step, per-kind sweep launches / time, witness check."""
import json
import sys

for f in sys.argv[1:]:
    try:
        with open(f) as fh:
            line = json.loads([x for x in fh.read().strip().splitlines() if x.startswith("{")][-1])
        kinds = {k: (v["launches"], round(v["ms"], 1)) for k, v in line["roofline"]["per_kind"].items()}
        wc = (line.get("witness_check") or {}).get("worst_setting_sigmas")
        print(f, f"{line['value']:.4g} {line['unit']}", f"{line['ms_per_step']:.1f} ms/step", kinds, "worst sigma", wc,
              "states", line.get("states_swept_per_step_rank0"), "e2e", (line.get("e2e") or {}).get("value"))
    except Exception as exc:       # noqa: BLE001
        print(f, "unreadable:", exc)
