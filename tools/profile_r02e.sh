#!/bin/bash
# DRAM traffic of every sweep launch (register-only sweeps and the shared passes of sibling groups) of the first three
# graphs of the headline job: one application-replay pass, three metrics.  Reduced by profiles/extract_r02aq.py.
TAG=${1:-r02aq}
OUT=gpurun_out
CMD="python bench.py --graphs 3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
timeout 200 $CMD > $OUT/${TAG}_bench_g3.json 2> $OUT/${TAG}_bench_g3.err; echo "bench rc=$?"
timeout 400 ncu --replay-mode application --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:"sweep_direct|leaf_cross" -c 2000 --csv --log-file $OUT/${TAG}_traffic_g3.csv $CMD > $OUT/${TAG}_ncu2.log 2>&1
echo "traffic rc=$?"
ls -la $OUT/${TAG}_*
