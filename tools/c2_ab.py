#!/usr/bin/env python
"""BASELINE.json configs[1] (C2) end to end through qcd.simulate(), with the stage breakdown bench.py reports.
    QCD_DM_ON_CHIP=0 python tools/c2_ab.py [n_circuits]     # lowered sweep route, for comparison"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
import qcdenoise_b200 as qcd  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
out = bench.c2_end_to_end(qcd, torch, n)
out["dm_on_chip"] = os.environ.get("QCD_DM_ON_CHIP", "1")
print(json.dumps(out))
