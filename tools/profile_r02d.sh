#!/bin/bash
# Launch list (one metric, one pass: no kernel replay) of the first three graphs of the headline job with sibling groups.
TAG=${1:-r02ao}
OUT=gpurun_out
CMD="python bench.py --graphs 3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
timeout 200 $CMD > $OUT/${TAG}_bench_g3.json 2> $OUT/${TAG}_bench_g3.err; echo "bench rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file $OUT/${TAG}_launches_g3.csv \
    $CMD > $OUT/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
python profiles/summarize_launches.py $OUT/${TAG}_launches_g3.csv
