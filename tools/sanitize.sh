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
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard|c2_full|c4_arm|Invalid|at 0x" | head -20
done
