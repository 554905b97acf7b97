import json, sys
import torch
sys.path.insert(0, ".")
from tools.microbench import bench_env
out = []
bench_env(out)
json.dump(out, open("gpurun_out/microbench_env.json", "w"), indent=1)
