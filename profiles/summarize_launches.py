#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (count, total ms, share)."""
import collections
import csv
import io
import sys


def load(path):
    txt = open(path).read()
    return list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))


def summarize(rows):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("qcd::<unnamed>::", "")[-70:]
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    out = ["| launches | total ms | share | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | `{k}` |")
    out.append(f"| {sum(v[0] for v in agg.values())} | {tot / 1e6:.3f} | 100% | all |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(load(sys.argv[1])))
