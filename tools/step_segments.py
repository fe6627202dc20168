"""In-situ durations of the fused step's segments inside the replayed CUDA graph (event-record graph nodes; warm L2,
back-to-back launches — what ncu's cold, serialised per-kernel times cannot show).
    python tools/step_segments.py [frames] [size]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harp_b200 import scene  # noqa: E402
from harp_b200.step import HarpFitter  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
hc, _, _ = scene.make_context(device="cuda:0", model="hand")
sc = scene.make_scene(hc, B, S, seed=1)
ft = HarpFitter(hc, sc["params"], sc["frames"], sc["ref_verts"], sc["uv_mask"], S, sc["focal"])
fid = torch.arange(B, device="cuda")
oa, on = sc["offsets"]["albedo"], sc["offsets"]["normal"]
hc.profile(2)
replay = ft.capture(fid, oa, on, stage="both", frame_range=(0, B), keep_profile=True)
for _ in range(5):
    replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 20
e0.record()
for _ in range(N):
    replay()
e1.record()
torch.cuda.synchronize()
seg = hc.profile_read(detailed=True)
out = {"frames": B, "size": S, "ms_per_step": e0.elapsed_time(e1) / N,
       "segments_ms_last_replay": {k: round(v[0] / max(v[1], 1), 4) for k, v in seg.items()}}
out["accounted_ms"] = round(sum(v for k, v in out["segments_ms_last_replay"].items()), 4)
print(json.dumps(out))
if os.environ.get("HARP_TIMELINE"):
    for name, t0, t1 in sorted(hc.profile_timeline(), key=lambda r: r[1]):
        print(f"  {name:14s} {1e3 * t0:8.1f} -> {1e3 * t1:8.1f} us  ({1e3 * (t1 - t0):6.1f})")
