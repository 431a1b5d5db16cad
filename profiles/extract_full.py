#!/usr/bin/env python
"""Reduce an `ncu --set full` capture of the sweep kernel to the per-launch figures quoted in DESIGN.md / bench.py.

    ncu -i gpurun_out/prof_direct_r1t.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/extract_full.py /tmp/raw.csv r01t

Writes profiles/<tag>_sweep_direct_ncu_full_extract.csv (selected raw metrics), <tag>_sweep_direct_launches.json
(per launch: states, time, DRAM bytes, algorithmic bytes) and <tag>_traffic.json (DRAM bytes per algorithmic byte over
the inner-level launches, where every state is read and written: the factor bench.py uses for roofline.traffic)."""
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic"]
STATE_BYTES = 2 ** 20 * 8          # complex64, n = 20
CTAS_PER_STATE = 32                # 2^20 amplitudes / (128 threads x 16 amplitudes x 16 groups per CTA)


def main(raw, tag):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(os.path.join(HERE, f"{tag}_sweep_direct_ncu_full_extract.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in data:
            w.writerow([r[i] for i in cols])
    g = lambda r, k: float(r[hdr.index(k)])       # noqa: E731
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    launches = []
    for r in data:
        name = r[hdr.index("Kernel Name")]
        m = re.search(r"sweep_direct_kernel<(\w+),\s*(\d+),\s*(\d+)>", name)
        rd = g(r, "dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = g(r, "dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
        ms = g(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
        states = g(r, "launch__grid_size") / CTAS_PER_STATE
        launches.append({"kernel": m.group(0) if m else name[:60], "kind": int(m.group(3)) if m else None,
                         "vec128": int(m.group(2)) if m else None, "states": states, "ms": ms,
                         "dram_read_GB": rd / 1e9, "dram_write_GB": wr / 1e9, "dram_TBps": (rd + wr) / ms / 1e9,
                         "algorithmic_GB_if_all_stored": 2 * states * STATE_BYTES / 1e9,
                         "algorithmic_TBps_if_all_stored": 2 * states * STATE_BYTES / ms / 1e9,
                         "issue_active_pct": g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "thread_inst_per_16_amplitudes": 32 * g(r, "smsp__inst_executed.sum") / (states * 2 ** 20 / 16),
                         "registers": g(r, "launch__registers_per_thread"),
                         "l2_hit_pct": g(r, "lts__t_sector_hit_rate.pct")})
    with open(os.path.join(HERE, f"{tag}_sweep_direct_launches.json"), "w") as f:
        json.dump(launches, f, indent=1)
    inner = [x for x in launches if x["kind"] == 1 and x["states"] >= 64]
    alg = sum(x["algorithmic_GB_if_all_stored"] for x in inner)
    dram = sum(x["dram_read_GB"] + x["dram_write_GB"] for x in inner)
    with open(os.path.join(HERE, f"{tag}_traffic.json"), "w") as f:
        json.dump({"source": f"profiles/{tag}_sweep_direct_launches.json: the {len(inner)} inner-level launches (every state "
                             "read and written) of the ncu --set full capture",
                   "dram_bytes_per_algorithmic_byte": dram / alg,
                   "note": "leaf-level launches skip the write of leaf states sampled by <= 8 trajectories"}, f, indent=1)
    for x in launches:
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in x.items()})
    print("dram/algorithmic over inner launches:", dram / alg)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
