#!/bin/bash
# DRAM traffic of every sweep launch of `bench.py --graphs 3` (graphs 0-2 of the job: 9, 7 and 17 noise sites -- the heavy
# graph dominates, as in the 8-graph headline job), application replay, own timeouts.  Reduced by profiles/extract_r02.py.
TAG=${1:-r02v}
OUT=gpurun_out
CMD="python bench.py --graphs 3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
timeout 200 $CMD > $OUT/${TAG}_bench_g1.json 2> $OUT/${TAG}_bench_g1.err; echo "bench rc=$?"
timeout 600 ncu --replay-mode application --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    -k regex:sweep_ --csv --log-file $OUT/${TAG}_traffic_g1.csv $CMD > $OUT/${TAG}_ncu2.log 2>&1; echo "traffic rc=$?"
ls -la $OUT/${TAG}_*
