import json, sys
import torch
sys.path.insert(0, ".")
from tools.microbench import bench_agent
out = []
bench_agent(out)
json.dump(out, open("gpurun_out/microbench_agent.json", "w"), indent=1)
