#!/bin/bash
# Round-2 measurement recipe (run on the GPU box through gpurun, ONE GPU):  bash tools/profile_r02.sh <tag>
#   1. the default bench line (no profiler)                                         -> gpurun_out/<tag>_bench.json
#   2. the profiled command WITHOUT ncu (one graph of the job, 1 warm-up + 1 step)  -> gpurun_out/<tag>_bench_g1.json
#   3. ncu launch list of the same command (gpu__time_duration, every kernel)      -> gpurun_out/<tag>_launches_g1.csv
#   4. DRAM traffic of EVERY sweep launch of the same command                       -> gpurun_out/<tag>_traffic_g1.csv
#   5. ncu --set full (+ source) of 12 sweep launches of the timed step            -> gpurun_out/<tag>_sweep_direct.ncu-rep
# Numbers printed under ncu are never bench values; profiles/extract_r02.py reduces 2-5 to the files under profiles/.
TAG=${1:-r02f}
OUT=gpurun_out
CMD="python bench.py --graphs 1 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
$CMD > $OUT/${TAG}_bench_g1.json 2> $OUT/${TAG}_bench_g1.err; rc=$?; echo "g1 rc=$rc"
[ $rc -ne 0 ] && exit $rc
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches_g1.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:sweep_ -c 1000 --csv --log-file $OUT/${TAG}_traffic_g1.csv $CMD > $OUT/${TAG}_ncu2.log 2>&1; echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sweep_direct -s 70 -c 12 -f -o $OUT/${TAG}_sweep_direct $CMD > $OUT/${TAG}_ncu3.log 2>&1; echo "full rc=$?"
ls -la $OUT/${TAG}_*
