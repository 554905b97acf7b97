#!/bin/bash
# Final evidence pass of round 2 (one gpurun call, one GPU).  Plain runs first (each must exit 0), then ncu of the same
# commands; reports are condensed to CSV on the box (the .ncu-rep files stay there: gpurun_out is limited to 64 MiB).
P="python tools/prof_cases.py"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r2t_gputests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2t_gputests.log
timeout 600 python bench.py > $O/r2t_bench.json 2> $O/r2t_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2t_bench_reference_arm.json 2> $O/r2t_bench_ref.err; echo "ref rc=$?"
# launch list of the bench's C2 part (rollout steps, host steps, C1-size learner)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2t_launches_bench_c2.csv python bench.py --steps 200 --warmup 5 --learner-steps 2 --no-cpu-baseline --only c2 > $O/ncu_list.log 2>&1
for c in rollout_c2 env_c3 env_c3_64k env_c2_1m agent_c3; do timeout 120 $P $c > $O/plain_$c.log 2>&1 || { echo "plain $c failed"; tail -5 $O/plain_$c.log; }; done
# the bench's timed launch: the persistent rollout kernel (third launch)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agent_forward_tc2 -s 2 -c 1 -o /tmp/r2t_rollout -f $P rollout_c2 > $O/ncu_rollout.log 2>&1
ncu -i /tmp/r2t_rollout.ncu-rep --page raw --csv > $O/r2t_rollout.raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:agent_forward_tc2 -s 2 -c 1 -o /tmp/r2t_agent_c3 -f $P agent_c3 > $O/ncu_agent_c3.log 2>&1
ncu -i /tmp/r2t_agent_c3.ncu-rep --page raw --csv > $O/r2t_agent_c3.raw.csv 2>/dev/null
for c in env_c3 env_c3_64k env_c2_1m; do
  timeout 200 ncu --set full --clock-control none -k regex:env_step2 -s 1 -c 1 -o /tmp/r2t_$c -f $P $c > $O/ncu_$c.log 2>&1
  ncu -i /tmp/r2t_$c.ncu-rep --page raw --csv > $O/r2t_$c.raw.csv 2>/dev/null
done
# learner at C4: launch list of the third train step, and a full capture of the recurrence launch
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "train2/" --csv --log-file $O/r2t_launches_learner_c4.csv $P learner_c4 3 > $O/ncu_c4_list.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gru_recurrence -s 2 -c 1 -o /tmp/r2t_rec -f $P learner_c4 3 > $O/ncu_rec.log 2>&1
ncu -i /tmp/r2t_rec.ncu-rep --page raw --csv > $O/r2t_rec.raw.csv 2>/dev/null
# phase stamps (profiling build of the library)
L=tools/_prof/libmacjd_prof.so
MACJD_LIB_PATH=$L timeout 60 python tools/rollout_phase_profile.py > $O/r2t_rollout_phase.txt 2>&1
MACJD_LIB_PATH=$L timeout 60 python tools/rec_phase_profile.py > $O/r2t_rec_phase.txt 2>&1
TC_O=176 TC_A=33 MACJD_LIB_PATH=$L timeout 60 python tools/tc_phase_profile.py 65536 3 1 > $O/r2t_agent_c3_phase.txt 2>&1
ls -la $O | tail -30
