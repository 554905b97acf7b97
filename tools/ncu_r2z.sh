#!/bin/bash
# Closing evidence of round 2 on the final code: GPU tests, bench lines of both arms, launch list of the bench command.
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r2z_gputests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2z_gputests.log
timeout 600 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/r2z_bench_reference_arm.json 2> $O/r2z_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2z_launches_bench_c2.csv python bench.py --only none --no-cpu-baseline --steps 100 --warmup 5 > $O/ncu_launches_bench.log 2>&1; echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
