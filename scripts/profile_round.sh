#!/bin/bash
# The ncu evidence of a round (run on the GPU box through gpurun; outputs under gpurun_out/, summarised into profiles/ by
# scripts/ncu_summary.py and scripts/kernel_times.py).   bash scripts/profile_round.sh r02
R=${1:-r02}
B="--no-cpu-baseline --no-other-workloads"
# 1. launch list of the bench command (serialised, cold-cache per-launch times: the SHARE of a kernel is what must agree)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_pbf130k.csv \
  python bench.py --steps 4 --warmup 3 $B > gpurun_out/${R}_launches.log 2>&1
# 2. every kernel of one 130k step (15 launches), step ~200 of the dam break
ncu --set full --clock-control none --import-source on -k regex:Kernel$ --launch-skip 3000 --launch-count 15 -f \
  -o gpurun_out/${R}_pbf130k_step python bench.py --steps 4 --warmup 200 $B > gpurun_out/${R}_full130k.log 2>&1
# 3. every kernel of one 16.7M step (16 launches: one more sort pass), third step; fewer sections: every replay pass
#    saves and restores GBs of state
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats \
  --section LaunchStats --section Occupancy --clock-control none -k regex:Kernel$ --launch-skip 32 --launch-count 16 -f \
  -o gpurun_out/${R}_pbf16m_step python scripts/step16m.py 4 > gpurun_out/${R}_full16m.log 2>&1
tail -n 2 gpurun_out/${R}_full130k.log; tail -n 2 gpurun_out/${R}_full16m.log
ls -la gpurun_out/${R}_*
