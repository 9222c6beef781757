#!/bin/bash
# Round profile pass (run under gpurun, one GPU): plain bench first, then the ncu launch list of the
# same command, one `--set full` capture of the posterior kernel, and a lighter capture (speed-of-light,
# compute, memory, warp-state sections; one launch per kernel) of every other kernel.  The reports are
# condensed on the box (profiles/summarize_ncu.py); only the posterior report travels back.
# Numbers printed by a run under ncu are never used as bench values.
set -u
TAG=${1:-r2}
OUT=gpurun_out
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --configs 2 > $OUT/${TAG}_ncu_launch.log 2>&1
python profiles/summarize_ncu.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.txt
ncu --set full --clock-control none --import-source on -k regex:bf_posterior -s 3 -c 1 -f -o $OUT/${TAG}_posterior \
    python bench.py --steps 2 --warmup 3 --configs 2 > $OUT/${TAG}_ncu_full.log 2>&1
python profiles/summarize_ncu.py full $OUT/${TAG}_posterior.ncu-rep > $OUT/${TAG}_posterior_ncu_full.txt
python tools/profile_secondary.py > $OUT/${TAG}_secondary.json 2> $OUT/${TAG}_secondary.err || exit 2
BF_PROFILE_ONCE=1 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis \
    --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none -k regex:^bf_ -c 90 \
    -f -o /tmp/${TAG}_secondary python tools/profile_secondary.py > $OUT/${TAG}_ncu_secondary.log 2>&1
python profiles/summarize_ncu.py full /tmp/${TAG}_secondary.ncu-rep > $OUT/${TAG}_secondary_ncu.txt
cat $OUT/${TAG}_secondary.json
ls -la $OUT/${TAG}_* /tmp/${TAG}_secondary.ncu-rep
