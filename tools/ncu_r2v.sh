#!/bin/bash
# Last evidence refresh of round 2 after the GEMM changes (pre-split weights, full-wave split-K, narrow outputs): GPU tests,
# the bench lines, the C4 launch list and a sectioned capture of the C4 GEMM / column-sum launches.
P="python tools/prof_cases.py"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r2v_gputests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2v_gputests.log
timeout 600 python bench.py > $O/r2v_bench.json 2> $O/r2v_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2v_bench_reference_arm.json 2> $O/r2v_bench_ref.err; echo "ref rc=$?"
timeout 120 $P learner_c4 3 > $O/plain_learner_c4.log 2>&1 || echo "plain learner_c4 failed"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "train2/" --csv --log-file $O/r2v_launches_learner_c4.csv $P learner_c4 3 > $O/ncu_c4_list.log 2>&1
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
timeout 500 ncu $SECS --clock-control none --nvtx --nvtx-include "train2/" -k regex:"tc_gemm|tc_pack|colsum_partial|sgemm" -o /tmp/r2v_gemm_c4 -f $P learner_c4 3 > $O/ncu_c4_gemm.log 2>&1
ncu -i /tmp/r2v_gemm_c4.ncu-rep --page raw --csv > $O/r2v_gemm_c4.raw.csv 2>/dev/null
ls -la $O | tail -12
