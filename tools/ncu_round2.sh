#!/bin/bash
# One gpurun call: plain runs first (each must exit 0), then the ncu captures of the same commands.
set -x
P="python tools/prof_cases.py"
NCU="ncu --set full --clock-control none --import-source on"
for c in learner_c1 learner_c4 agent_c3 env_c3 env_c3_64k; do $P $c > gpurun_out/plain_$c.log 2>&1 || exit 1; done
# learner, C1 size: every kernel of the 3rd train step (skip two steps' worth of launches by nvtx range)
$NCU --nvtx --nvtx-include "train2/" -o gpurun_out/r2_learner_c1 -f $P learner_c1 > gpurun_out/ncu_learner_c1.log 2>&1
# learner, C4 size: the learner kernels proper (the 100-step unroll kernels are captured at C1 size above)
$NCU --nvtx --nvtx-include "train2/" -k regex:'mix_|sgemm|splitk|colsum|td_|adam|sumsq|clip_coef|qhead|layernorm|gather_q|replay' \
     -o gpurun_out/r2_learner_c4 -f $P learner_c4 > gpurun_out/ncu_learner_c4.log 2>&1
$NCU -k regex:agent_forward_tc2 -s 2 -c 1 -o gpurun_out/r2_agent_c3 -f $P agent_c3 > gpurun_out/ncu_agent_c3.log 2>&1
$NCU -k regex:env_step -s 3 -c 1 -o gpurun_out/r2_env_c3 -f $P env_c3 > gpurun_out/ncu_env_c3.log 2>&1
$NCU -k regex:env_step -s 3 -c 1 -o gpurun_out/r2_env_c3_64k -f $P env_c3_64k > gpurun_out/ncu_env_c3_64k.log 2>&1
ls -la gpurun_out/*.ncu-rep
