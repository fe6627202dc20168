"""Where the Level-1 drop-in loop body spends its time (torch profiler, CUDA time by op / kernel).
    python tools/profile_dropin.py [frames] [size]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from harp_b200 import scene  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
hc, _, _ = scene.make_context(device="cuda:0", model="hand")
sc = scene.make_scene(hc, B, S, seed=1)
print(bench.dropin_leg(sc, hc, B, S, steps=3, warmup=2))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    bench.dropin_leg(sc, hc, B, S, steps=2, warmup=0)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print("\n==== by host time (the step is bound by the host: CPU RNG draws of the texture regularisers, pageable host->device copies, ATen launches)")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
