#!/usr/bin/env python
"""Reduce the round-2 ncu captures (tools/profile_r02.sh) to the files bench.py and DESIGN.md quote.

    python profiles/extract_r02.py gpurun_out r02f

Inputs (gpurun_out/<tag>_*):
    _bench_g1.json     bench line of the profiled command run WITHOUT ncu (algorithmic bytes per launch kind per step)
    _launches_g1.csv   ncu --metrics gpu__time_duration.sum of every kernel of that command
    _traffic_g1.csv    ncu DRAM bytes / time / issue-slot metrics of EVERY sweep launch of that command
Outputs (profiles/):
    <tag>_launch_summary.md        kernels by share of GPU time (the sweep kernels' share must agree with bench.py's
                                   `sweep_share_of_step`; ncu per-launch times are cold-cache and serialised, the SHARE is
                                   what is compared)
    <tag>_sweep_launches.json      per sweep launch: kernel, kind, states, time, DRAM bytes, DRAM GB/s, issue-slot %
    r02_traffic.json               DRAM bytes per ALGORITHMIC byte per launch kind (inner / leaf / mixed) -- what bench.py
                                   turns into roofline.traffic and roofline.dram_frac; the capture holds the warm-up step
                                   and the timed step (identical work), so each kind's DRAM bytes are divided by twice
                                   the per-step algorithmic bytes of the bench line
"""
import collections
import csv
import io
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from summarize_launches import load, summarize  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0, "": 1.0,
        "inst": 1.0}
KIND = {0: "plain", 1: "inner", 2: "leaf", 3: "mixed"}
STATE_BYTES = 2 ** 20 * 8


def long_to_launches(path):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    out = collections.OrderedDict()
    for r in rows:
        d = out.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        except ValueError:
            pass
    return list(out.values())


def main(src, tag):
    with open(os.path.join(src, f"{tag}_bench_g1.json")) as f:
        line = json.loads([l for l in f if l.startswith("{")][-1])
    steps = line["steps"] + max(line["warmup"], 1)
    kinds = line["roofline"]["per_kind"]
    graphs = line["config"]["graphs_per_step"]
    cmd = f"python bench.py --graphs {graphs} --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
    lpath = os.path.join(src, f"{tag}_launches_g1.csv")
    if os.path.exists(lpath):
        with open(os.path.join(HERE, f"{tag}_launch_summary.md"), "w") as f:
            rows = load(lpath)
            f.write(f"# {tag}: `{cmd}` under `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
            f.write(summarize(rows) + "\n\n")
            sweep = sum(float(r["Metric Value"]) for r in rows if "sweep_" in r["Kernel Name"])
            tot = sum(float(r["Metric Value"]) for r in rows if "at::" not in r["Kernel Name"])
            f.write(f"sweep kernels: {100 * sweep / tot:.1f} % of the engine's kernel time under ncu; bench.py (CUDA events, no "
                    f"profiler) reports sweep_share_of_step = {100 * line['roofline']['sweep_share_of_step']:.1f} %\n")
    launches = []
    for d in long_to_launches(os.path.join(src, f"{tag}_traffic_g1.csv")):
        m = re.search(r"sweep_direct_kernel<(\w+),\s*\(?\w*\)?(\d+),\s*(\d+)>", d["kernel"])
        ms = d.get("gpu__time_duration.sum", 0.0)
        dram = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        ctas = float(d.get("launch__grid_size", 0))
        launches.append({"kernel": m.group(0) if m else d["kernel"].split("(")[0][-48:],
                         "kind": KIND[int(m.group(3))] if m else "tile",
                         "ctas": ctas, "ms": ms, "dram_read_GB": d.get("dram__bytes_read.sum", 0.0) / 1e9,
                         "dram_write_GB": d.get("dram__bytes_write.sum", 0.0) / 1e9,
                         "dram_GBps": dram / ms / 1e6 if ms else None,
                         "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "l2_hit_pct": d.get("lts__t_sector_hit_rate.pct"),
                         "warp_inst": d.get("smsp__inst_executed.sum")})
    with open(os.path.join(HERE, f"{tag}_sweep_launches.json"), "w") as f:
        json.dump(launches, f, indent=0)
    ratio, detail = {}, {}
    for nm in ("inner", "leaf", "mixed"):
        sel = [x for x in launches if x["kind"] == nm]
        if not sel or nm not in kinds:
            continue
        dram = sum(x["dram_read_GB"] + x["dram_write_GB"] for x in sel) * 1e9
        ms = sum(x["ms"] for x in sel)
        alg = kinds[nm]["algorithmic_GBps"] * 1e9 * kinds[nm]["ms"] * 1e-3 / line["steps"] * steps
        ratio[nm] = dram / alg
        detail[nm] = {"launches_captured": len(sel), "launches_per_step_bench": kinds[nm]["launches"] / line["steps"],
                      "dram_GB": dram / 1e9, "algorithmic_GB": alg / 1e9, "ncu_ms": ms,
                      "dram_GBps_under_ncu": dram / ms / 1e6,
                      "issue_active_pct_time_weighted": sum(x["issue_active_pct"] * x["ms"] for x in sel) / ms,
                      "l2_hit_pct_time_weighted": sum(x["l2_hit_pct"] * x["ms"] for x in sel) / ms}
    with open(os.path.join(HERE, "r02_traffic.json"), "w") as f:
        json.dump({"source": f"profiles/{tag}_sweep_launches.json: every sweep launch of `{cmd}` (warm-up + timed step; noise "
                             f"sites per graph {line['config']['noise_sites_per_graph']}) under ncu (application replay), "
                             "dram__bytes_read.sum + dram__bytes_write.sum",
                   "dram_bytes_per_algorithmic_byte": ratio, "per_kind": detail}, f, indent=1)
    print(json.dumps({"ratio": ratio, "detail": detail}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
