#!/bin/bash
# Evidence pass r2x (last session of round 2): GPU tests, bench lines of both arms, launch lists of the bench command and of
# the C1 learner step, ncu --set full of the learner kernels at C1 (incl. the row-split recurrence kernel).
P="python tools/prof_cases.py"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r2x_gputests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2x_gputests.log
timeout 600 python bench.py > $O/r2x_bench.json 2> $O/r2x_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/r2x_bench_reference_arm.json 2> $O/r2x_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2x_launches_bench_c2.csv python bench.py --only none --no-cpu-baseline --steps 100 --warmup 5 > $O/ncu_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 120 $P learner_c1 > $O/plain_learner_c1.log 2>&1 || echo "plain learner_c1 failed"
timeout 400 ncu --set full --clock-control none -k regex:'macjd|tc::' -c 130 -o /tmp/r2x_learner_c1 -f $P learner_c1 2 > $O/ncu_learner_c1.log 2>&1; echo "ncu learner rc=$?"
ncu -i /tmp/r2x_learner_c1.ncu-rep --page raw --csv > $O/r2x_learner_c1.raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r2x_learner_c1.raw.csv $O/r2x_learner_c1.summary.csv --per-kernel
ls -la $O | tail -14
