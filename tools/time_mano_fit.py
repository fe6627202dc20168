"""Wall time of the upstream per-frame MANO fit (500 + 700 Adam steps, hand_utils.py:16-133) on the B200 path.
    python tools/time_mano_fit.py [frames]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harp_b200 import mano_fit, synth  # noqa: E402
from harp_b200.manolayer import ManoLayer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tmpl = synth.load_template("hand")
pack = synth.make_mano_pack(seed=0, template=tmpl)
layer = ManoLayer(pack, tmpl["faces_uvs"], tmpl["verts_uvs"])
seq = synth.make_sequence(B, seed=1)
trans = torch.tensor([[0.0, 0.0, 0.5]]).repeat(B, 1)
with torch.no_grad():
    verts_mm, _ = layer(torch.cat((seq["rot"], seq["pose"]), 1).cuda(), seq["shape"][None].repeat(B, 1).cuda(), trans.cuda())
pred = verts_mm / 1000.0 + 0.0005 * torch.randn_like(verts_mm)
mano_fit.optimize_for_mano_param(pred, layer, epoch_coarse=5, epoch_fine=5)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
out = mano_fit.optimize_for_mano_param(pred, layer)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"what": "optimize_for_mano_param, 500 + 700 Adam steps, CUDA-graph replay", "frames": B, "seconds": dt,
                  "tries": mano_fit.optimize_for_mano_param.last_tries, "final_mse_mm2": mano_fit.optimize_for_mano_param.last_loss,
                  "us_per_iteration": dt / (1200 * mano_fit.optimize_for_mano_param.last_tries) * 1e6}))
