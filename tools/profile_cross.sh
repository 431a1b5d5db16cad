#!/bin/bash
# ncu sections of the leaf_cross kernel (shared passes of sibling groups) on a one-graph job, application replay (kernel
# replay has to save and restore the state arena: it never finished on this workload), own timeouts.
TAG=${1:-r02aj}
OUT=gpurun_out
CMD="python bench.py --graphs 1 --steps 1 --warmup 1 --shots 300000 --no-cpu --no-e2e --no-other"
timeout 120 $CMD > $OUT/${TAG}_bench_g1.json 2> $OUT/${TAG}_bench_g1.err; echo "bench rc=$?"
timeout 500 ncu --replay-mode application --clock-control none --import-source on \
    --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis \
    --section Occupancy --section LaunchStats --section InstructionStats --section SourceCounters \
    -k regex:leaf_cross -c 3 -o $OUT/${TAG}_leaf_cross $CMD > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/${TAG}_ncu.log
ls -la $OUT/${TAG}_*
