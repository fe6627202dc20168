#!/bin/bash
# compute-sanitizer over one small fused step + the stand-alone render ops (run on the GPU box):
#   gpurun -- bash tools/sanitize.sh
set -u
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
from util import load_golden, golden_case
import test_gpu_kernels as T
for name in ("c2_full_128_sub", "c4_arm_128_sub"):
    g = load_golden(name)
    case = golden_case(g)
    model, topo, params, batch, cfg = case
    ft, hc = T._fitter_from_case(case)
    B = params["pose"].shape[0]
    fid = torch.arange(B, device="cuda")
    for _ in range(2):
        ft.step(fid, batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous(), stage="both", update=True)
    torch.cuda.synchronize()
    print(name, ft.loss_dict()["total"], hc.counters())
# the Level-1 drop-in loop body (stand-alone silhouette / shade ops forward and backward, single-term mesh losses, hit mask,
# evaluation renderers)
import test_gpu_dropin as D
for name in ("c2_full_128_sub", "c4_arm_128_sub"):
    loss, total, leaf, y_sil_pred, y_pred = D._loop_body(load_golden(name), "both")
    torch.cuda.synchronize()
    print("dropin", name, float(total))
# upstream per-frame MANO fit (a few eager iterations)
import numpy as np
from harp_b200 import mano_fit, synth
from harp_b200.manolayer import ManoLayer
tmpl = synth.load_template("hand"); pack = synth.make_mano_pack(seed=0, template=tmpl)
layer = ManoLayer(pack, tmpl["faces_uvs"], tmpl["verts_uvs"])
g = load_golden("fit_mano_b2")
mano_fit.optimize_for_mano_param(torch.tensor(g["pred_vertices"]), layer, epoch_coarse=3, epoch_fine=3, use_graph=False)
print("mano_fit", mano_fit.optimize_for_mano_param.last_loss)
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --print-limit 100 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard|c2_full|c4_arm|mano_fit|dropin|Invalid|at 0x" | sort | uniq -c | sort -rn | head -30
done
