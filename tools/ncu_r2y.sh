#!/bin/bash
# Evidence pass r2y: ncu --set full of every kernel of ONE eager C1 train step (B = 32 x T = 100; the NVTX range of the second
# step) on the final learner path (row-split recurrence kernel, one-launch Q-head re-pack), plus its launch list.
P="python tools/prof_cases.py"
O=gpurun_out
timeout 120 $P learner_c1 2 > $O/plain_learner_c1.log 2>&1 || echo "plain learner_c1 failed"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "train1/" --csv --log-file $O/r2y_launches_learner_c1.csv $P learner_c1 2 > $O/ncu_c1_list.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --nvtx --nvtx-include "train1/" -o /tmp/r2y_learner_c1 -f $P learner_c1 2 > $O/ncu_learner_c1.log 2>&1; echo "ncu learner rc=$?"
ncu -i /tmp/r2y_learner_c1.ncu-rep --page raw --csv > $O/r2y_learner_c1.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2y_learner_c1.raw.csv $O/r2y_learner_c1.summary.csv --per-kernel
wc -l $O/r2y_learner_c1.summary.csv $O/r2y_launches_learner_c1.csv
