"""GPU time of the C1 train step (B = 32 x T = 100) through QMixLearner.train_sampled (the replayed step graph) under the
current environment switches -- run once per setting to A/B a kernel choice on one box:
   MACJD_TC_GEMM=0 python tools/ab_learner_c1.py        (every learner GEMM on the FP32 kernel)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.test_gpu_learner import _sampled_learner   # noqa: E402

learner, buf = _sampled_learner(seed=5, n_envs=256)
learner.args.target_update_interval = 200
np.random.seed(1)
for _ in range(10):
    learner.train_sampled(buf, 32, {})
torch.cuda.synchronize()
best = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        s = learner.train_sampled(buf, 32, {})
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 100)
print({k: os.environ.get(k) for k in ("MACJD_TC_GEMM", "MACJD_TC_GEMM_MIN_LOG2", "MACJD_TRAIN_GRAPH")},
      "ms per step:", " ".join(f"{b:.4f}" for b in best), "stats", [round(float(x), 6) for x in s["stats_tensor"]])
