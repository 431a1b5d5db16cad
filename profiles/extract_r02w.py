#!/usr/bin/env python
"""DRAM bytes per algorithmic byte of the sweep launches of a HEAVY graph (graph 2 of the job, 17 noise sites), from a
single-pass application-replay capture of the timed step of `bench.py --graphs 3 --steps 1 --warmup 1` (the queue runs the
heaviest graph first: `-s 100 -c 64` covers its launches):

    ncu --replay-mode application --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        -k regex:sweep_direct -s 100 -c 64 --csv --log-file gpurun_out/r02w_traffic_heavy.csv python bench.py --graphs 3 ...
    python profiles/extract_r02w.py gpurun_out/r02w_traffic_heavy.csv

Heavy graphs dominate the 8-graph headline job, so these ratios (not graph 0's, profiles/r02m_traffic_graph0.json) are what
bench.py applies.  Inner launches: algorithmic bytes = states x 2 x 8 MiB exactly (states = CTAs / 32).  Leaf launches: an
unstored leaf is a read of its parent (8 MiB algorithmic), a stored one adds the write; the stored fraction is taken from
the DRAM writes themselves (writes / 8 MiB per leaf), so the leaf ratio is (reads + writes) / (1 + stored fraction)."""
import collections
import csv
import io
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "msecond": 1, "usecond": 1e-3,
     "nsecond": 1e-6}
STATE = 8 * 2 ** 20


def main(path):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r["ID"], {"k": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * U.get(r["Metric Unit"], 1)
    out, tot = [], {}
    for i, d in launches.items():
        m = re.search(r"sweep_direct_kernel<(\w+),\s*\(?\w*\)?(\d+),\s*(\d+)>", d["k"])
        kind = {0: "plain", 1: "inner", 2: "leaf", 3: "mixed"}[int(m.group(3))]
        states = int(re.match(r"\((\d+)", d["grid"]).group(1)) / 32
        ms, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        out.append({"id": int(i), "kind": kind, "states": states, "ms": ms, "dram_read_GB": rd / 1e9,
                    "dram_write_GB": wr / 1e9, "dram_TBps": (rd + wr) / ms / 1e9})
        t = tot.setdefault(kind, [0, 0.0, 0.0, 0.0, 0.0])
        t[0] += 1; t[1] += states; t[2] += ms; t[3] += rd; t[4] += wr
    with open(os.path.join(HERE, "r02w_sweep_launches_heavy_graph.json"), "w") as f:
        json.dump(out, f, indent=0)
    n_i, s_i, ms_i, rd_i, wr_i = tot["inner"]
    n_l, s_l, ms_l, rd_l, wr_l = tot["leaf"]
    stored = wr_l / (s_l * STATE)
    ratio = {"inner": (rd_i + wr_i) / (s_i * 2 * STATE), "leaf": (rd_l + wr_l) / (s_l * STATE * (1 + stored))}
    detail = {"inner": {"launches": n_i, "states": s_i, "ncu_ms": ms_i, "dram_GB": (rd_i + wr_i) / 1e9,
                        "dram_TBps_under_ncu": (rd_i + wr_i) / ms_i / 1e9, "dram_read_per_state_read": rd_i / (s_i * STATE),
                        "dram_write_per_state_written": wr_i / (s_i * STATE)},
              "leaf": {"launches": n_l, "states": s_l, "ncu_ms": ms_l, "dram_GB": (rd_l + wr_l) / 1e9,
                       "dram_TBps_under_ncu": (rd_l + wr_l) / ms_l / 1e9, "dram_read_per_leaf": rd_l / (s_l * STATE),
                       "stored_fraction_from_writes": stored, "us_per_leaf": 1e3 * ms_l / s_l}}
    with open(os.path.join(HERE, "r02_traffic.json"), "w") as f:
        json.dump({"source": "profiles/r02w_sweep_launches_heavy_graph.json: the 64 sweep launches of graph 2 (17 noise sites) in "
                             "the timed step of `bench.py --graphs 3 --steps 1 --warmup 1`, ncu application replay, "
                             "dram__bytes_read.sum + dram__bytes_write.sum; graph 0 alone (9 sites): "
                             "profiles/r02m_traffic_graph0.json (inner 0.79, leaf 0.37)",
                   "dram_bytes_per_algorithmic_byte": ratio, "per_kind": detail}, f, indent=1)
    print(json.dumps({"ratio": ratio, "detail": detail}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
