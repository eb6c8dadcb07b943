"""One training step at the bench shape (16 pairs x 2048 pts) for ncu launch lists:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_profile.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import se3_equi_graph_registration_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
r = bench.train_step_bench(P, dev, 0, 1, torch.cuda.synchronize, steps=steps, warmup=1)
print(r)
