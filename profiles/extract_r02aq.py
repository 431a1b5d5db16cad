#!/usr/bin/env python
"""DRAM bytes per algorithmic byte per launch kind with sibling groups at the leaf level (tools/profile_r02e.sh):

    python profiles/extract_r02aq.py gpurun_out r02aq

Inputs: <tag>_bench_g3.json (the profiled command WITHOUT ncu: algorithmic bytes per launch kind of the timed step) and
<tag>_traffic_g3.csv (ncu application replay: time, DRAM reads and writes of every sweep_direct / leaf_cross launch of the
warm-up step and the timed step -- identical work, so each kind's DRAM bytes are divided by twice the bench line's bytes).
inner = sweep_direct_kernel<., ., 1>; leaf = sweep_direct_kernel<., ., 2> (leaves that keep their own sweep) +
leaf_cross_kernel (shared passes: algorithmic bytes = one read of the parent per pass).
Outputs: profiles/<tag>_sweep_launches.json and profiles/r02_traffic.json (what bench.py applies)."""
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


def main(src, tag):
    with open(os.path.join(src, f"{tag}_bench_g3.json")) as f:
        line = json.loads([x for x in f if x.startswith("{")][-1])
    kinds = line["roofline"]["per_kind"]
    txt = open(os.path.join(src, f"{tag}_traffic_g3.csv")).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r["ID"], {"k": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * U.get(r["Metric Unit"], 1)
    out = []
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])       # launches, CTAs, ms, read, write
    for i, d in launches.items():
        m = re.search(r"sweep_direct_kernel<(\w+),\s*\(?\w*\)?(\d+),\s*(\d+)>", d["k"])
        if m:
            name = {0: "prep", 1: "inner", 2: "leaf_own", 3: "mixed"}[int(m.group(3))]
        elif "leaf_cross_kernel" in d["k"]:
            name = "leaf_shared"
        else:
            continue
        ctas = int(re.match(r"\((\d+)", d["grid"]).group(1))
        ms, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        out.append({"id": int(i), "kernel": name, "ctas": ctas, "ms": ms, "dram_read_GB": rd / 1e9, "dram_write_GB": wr / 1e9,
                    "dram_TBps": (rd + wr) / ms / 1e9 if ms else None})
        t = tot[name]
        t[0] += 1; t[1] += ctas; t[2] += ms; t[3] += rd; t[4] += wr
    with open(os.path.join(HERE, f"{tag}_sweep_launches.json"), "w") as f:
        json.dump(out, f, indent=0)
    alg = {k: v["algorithmic_GBps"] * 1e9 * v["ms"] * 1e-3 for k, v in kinds.items()}     # bytes of the timed step
    dram = {"inner": tot["inner"][3] + tot["inner"][4],
            "leaf": tot["leaf_own"][3] + tot["leaf_own"][4] + tot["leaf_shared"][3] + tot["leaf_shared"][4]}
    ratio = {k: dram[k] / (2 * alg[k]) for k in dram if alg.get(k)}
    detail = {}
    for name, (n, ctas, ms, rd, wr) in tot.items():
        per = 32.0                                                   # CTAs per state / per shared pass at n = 20
        detail[name] = {"launches": n, "states_or_passes": ctas / per, "ncu_ms": ms, "dram_GB": (rd + wr) / 1e9,
                        "dram_TBps_under_ncu": (rd + wr) / ms / 1e9 if ms else None,
                        "us_per_state_or_pass": 1e3 * ms / (ctas / per) if ctas else None,
                        "dram_MB_per_state_or_pass": (rd + wr) / (ctas / per) / 1e6 if ctas else None}
    with open(os.path.join(HERE, "r02_traffic.json"), "w") as f:
        json.dump({"source": f"profiles/{tag}_sweep_launches.json: every sweep_direct / leaf_cross launch of `python bench.py "
                             "--graphs 3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other` (graphs of 9, 7 and 17 noise sites; "
                             "warm-up step + timed step), ncu application replay, dram__bytes_read.sum + dram__bytes_write.sum, "
                             f"divided by twice the algorithmic bytes per kind of profiles/{tag}_bench_g3.json; before sibling "
                             "groups: inner 0.86, leaf 0.15 (profiles/r02w_sweep_launches_heavy_graph.json)",
                   "dram_bytes_per_algorithmic_byte": ratio, "per_kernel": detail,
                   "algorithmic_GB_timed_step": {k: v / 1e9 for k, v in alg.items()}}, f, indent=1)
    print(json.dumps({"ratio": ratio, "detail": detail}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
