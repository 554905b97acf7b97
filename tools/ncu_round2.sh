#!/bin/bash
# One gpurun call: plain runs first (each must exit 0), then ncu captures of the same commands.  Reports are turned
# into CSV on the box and the .ncu-rep files are NOT brought back (gpurun_out is limited to 64 MiB).
P="python tools/prof_cases.py"
O=gpurun_out
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
for c in learner_c1 agent_c3; do $P $c > $O/plain_$c.log 2>&1 || { echo "plain $c failed"; exit 1; }; done
# learner at C1 size: every kernel of the 3rd train step
ncu $SECS --clock-control none --nvtx --nvtx-include "train2/" -o /tmp/r2_learner_c1 -f $P learner_c1 > $O/ncu_learner_c1.log 2>&1
ncu -i /tmp/r2_learner_c1.ncu-rep --page raw --csv > $O/r2_learner_c1.raw.csv 2>/dev/null
# the C3 agent kernel, full set with source
ncu --set full --clock-control none --import-source on -k regex:agent_forward_tc2 -s 2 -c 1 -o /tmp/r2_agent_c3 -f $P agent_c3 > $O/ncu_agent_c3.log 2>&1
ncu -i /tmp/r2_agent_c3.ncu-rep --page raw --csv > $O/r2_agent_c3.raw.csv 2>/dev/null
ncu -i /tmp/r2_agent_c3.ncu-rep --page source --csv > $O/r2_agent_c3.source.csv 2>/dev/null
ls -la /tmp/*.ncu-rep; ls -la $O | tail -8
