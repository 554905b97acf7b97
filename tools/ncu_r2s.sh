#!/bin/bash
# Round-2 evidence pass (one gpurun call, one GPU): plain runs first (each must exit 0), then ncu captures of the same
# commands.  Reports are condensed to CSV on the box; the .ncu-rep files stay there (gpurun_out is limited to 64 MiB).
P="python tools/prof_cases.py"
O=gpurun_out
for c in rollout_c2 env_c3 env_c3_64k env_c2_1m; do $P $c > $O/plain_$c.log 2>&1 || { echo "plain $c failed"; tail -5 $O/plain_$c.log; exit 1; }; done
$P learner_c4 3 > $O/plain_learner_c4.log 2>&1 || { echo "plain learner_c4 failed"; exit 1; }
# 1. the bench's timed launch: the persistent rollout kernel (third launch: warm instruction cache, as in the bench's timed region)
ncu --set full --clock-control none --import-source on -k regex:agent_forward_tc2 -s 2 -c 1 -o /tmp/r2s_rollout -f $P rollout_c2 > $O/ncu_rollout.log 2>&1
ncu -i /tmp/r2s_rollout.ncu-rep --page raw --csv > $O/r2s_rollout.raw.csv 2>/dev/null
ncu -i /tmp/r2s_rollout.ncu-rep --page source --csv > $O/r2s_rollout.source.csv 2>/dev/null
# 2. env step kernel on derived tables at C3 (8 192 and 65 536 envs) and C2 at 1 M envs
for c in env_c3 env_c3_64k env_c2_1m; do
  ncu --set full --clock-control none -k regex:env_step2 -s 1 -c 1 -o /tmp/r2s_$c -f $P $c > $O/ncu_$c.log 2>&1
  ncu -i /tmp/r2s_$c.ncu-rep --page raw --csv > $O/r2s_$c.raw.csv 2>/dev/null
done
# 3. learner at C4: launch list of the third train step, then a sectioned capture of every kernel of that step
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "train2/" --csv --log-file $O/r2s_launches_learner_c4.csv $P learner_c4 3 > $O/ncu_c4_list.log 2>&1
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
ncu $SECS --clock-control none --nvtx --nvtx-include "train2/" -o /tmp/r2s_learner_c4 -f $P learner_c4 3 > $O/ncu_c4_full.log 2>&1
ncu -i /tmp/r2s_learner_c4.ncu-rep --page raw --csv > $O/r2s_learner_c4.raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep; ls -la $O | tail -12
