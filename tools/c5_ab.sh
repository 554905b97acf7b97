# A/B of the end-to-end loop (bench C5) on one box: the rollout's launches issued one by one vs replayed as one CUDA graph
for rg in 0 1 0 1; do
MACJD_BENCH_C5_ROLLOUT_GRAPH=$rg timeout 120 python bench.py --only c5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
c=json.loads(sys.stdin.read().strip().splitlines()[-1])['configs']['C5']
print('rollout_graph=$rg', round(c['env_agent_steps_per_sec']/1e6,2), 'M env-agent steps/s', round(c['train_episodes_per_sec']), 'train episodes/s, wall', round(c['wall_s'],3), 'loop', round(c['loop_only']['wall_s'],3))"
done
