#!/bin/bash
# Closing pass of round 2 on the final code: GPU tests, bench lines, the timed launch and the bench-size env kernel under ncu.
P="python tools/prof_cases.py"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r2w_gputests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2w_gputests.log
timeout 600 python bench.py > $O/r2w_bench.json 2> $O/r2w_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2w_bench_reference_arm.json 2> $O/r2w_bench_ref.err; echo "ref rc=$?"
for c in rollout_c2 env_c2; do timeout 120 $P $c > $O/plain_$c.log 2>&1 || echo "plain $c failed"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agent_forward_tc2 -s 2 -c 1 -o /tmp/r2w_rollout -f $P rollout_c2 > $O/ncu_rollout.log 2>&1
ncu -i /tmp/r2w_rollout.ncu-rep --page raw --csv > $O/r2w_rollout.raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:env_step2 -s 1 -c 1 -o /tmp/r2w_env_c2 -f $P env_c2 > $O/ncu_env_c2.log 2>&1
ncu -i /tmp/r2w_env_c2.ncu-rep --page raw --csv > $O/r2w_env_c2.raw.csv 2>/dev/null
MACJD_LIB_PATH=tools/_prof/libmacjd_prof.so timeout 60 python tools/rollout_phase_profile.py > $O/r2w_rollout_phase.txt 2>&1
ls -la $O | tail -12
