#!/bin/bash
# DRAM traffic + full sections of the sweep launches of `bench.py --graphs 1` (graph 0 of the job, 1 warm-up + 1 timed step)
# at the default 1 M shots.  Kernel replay would have to save and restore the tens of GB of state arena each launch writes
# (the round's first attempt did not finish in 23 minutes), so both captures use APPLICATION replay: the program is re-run
# once per metric pass, the launches are deterministic.  Every command carries its own timeout.
TAG=${1:-r02m}
OUT=gpurun_out
CMD="python bench.py --graphs 1 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
timeout 120 $CMD > $OUT/${TAG}_bench_g1.json 2> $OUT/${TAG}_bench_g1.err; echo "g1 rc=$?"
timeout 420 ncu --replay-mode application --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    -k regex:sweep_ --csv --log-file $OUT/${TAG}_traffic_g1.csv $CMD > $OUT/${TAG}_ncu2.log 2>&1; echo "traffic rc=$?"
timeout 900 ncu --replay-mode application --clock-control none --set detailed --import-source on -k regex:sweep_direct -s 11 -c 11 \
    -f -o $OUT/${TAG}_sweep_direct $CMD > $OUT/${TAG}_ncu3.log 2>&1; echo "detailed rc=$?"
ls -la $OUT/${TAG}_*
