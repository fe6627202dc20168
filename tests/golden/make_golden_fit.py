"""Golden vector for the upstream per-frame MANO fit: runs the REFERENCE'S OWN, UNMODIFIED
`metro_modifications/hand_utils.py:optimize_for_mano_param` (16-133) on CPU and stores its inputs and outputs.

Build-container only (needs /root/reference).  Two environment shims, neither a change of the reference's arithmetic:
  * `hand_utils.py` imports matplotlib, mpl_toolkits, trimesh and imageio at module level (plots / debug meshes only; not
    installed here): empty placeholder modules are injected so that the import succeeds; the fit never touches them
    (`SAVE_MESH_FOR_DEBUG = False`, line 26);
  * the function hard-codes `device = "cuda"` (49); there is no GPU in this container, so `.to("cuda")` on tensors and
    modules is mapped to the CPU.

    python tests/golden/make_golden_fit.py        # ~1 min of CPU
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "metro_modifications"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "trimesh", "imageio"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

_t_to, _m_to = torch.Tensor.to, torch.nn.Module.to


def _cpu(args):
    return tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)


torch.Tensor.to = lambda self, *a, **k: _t_to(self, *_cpu(a), **k)
torch.nn.Module.to = lambda self, *a, **k: _m_to(self, *_cpu(a), **k)

from harp_b200 import synth  # noqa: E402
from make_golden import ref_mano_layer  # noqa: E402  (the reference's ManoLayer on the synthetic MANO-shaped pack)


def main():
    import hand_utils  # the reference module, unmodified
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    tmpl = synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    layer = ref_mano_layer(pack)
    B = 2
    seq = synth.make_sequence(8, seed=1)
    rot, pose = seq["rot"][[1, 5]], seq["pose"][[1, 5]]
    shape = seq["shape"][None].repeat(B, 1)
    trans = torch.tensor([[0.02, -0.01, 0.45], [-0.03, 0.02, 0.50]])
    with torch.no_grad():
        verts_mm, _ = layer(torch.cat((rot, pose), 1), shape, trans)
    g = torch.Generator().manual_seed(7)
    pred_vertices = verts_mm / 1000.0 + 0.0005 * torch.randn(verts_mm.shape, generator=g)  # METRO-like: metres, 0.5 mm noise
    import contextlib
    import io
    import re
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):  # the function reports its stage losses only by print (91, 107)
        out = hand_utils.optimize_for_mano_param(pred_vertices.clone(), layer)
    log = buf.getvalue()
    print(log, end="")
    coarse = [float(x) for x in re.findall(r"After coarse alignment: ([0-9.]+)", log)]
    fine = [float(x) for x in re.findall(r"After fine alignment: ([0-9.]+)", log)]
    final = float(torch.nn.MSELoss()(torch.tensor(out["verts"]), pred_vertices * 1000.0))
    print("reference fit: final MSE (mm^2)", final)
    np.savez_compressed(os.path.join(HERE, "fit_mano_b2.npz"), pred_vertices=pred_vertices.numpy(),
                        gt_rot=rot.numpy(), gt_pose=pose.numpy(), gt_shape=shape.numpy(), gt_trans=trans.numpy(),
                        final_mse=np.float32(final), printed_coarse_loss=np.float32(coarse[-1]),
                        printed_fine_loss=np.float32(fine[-1]), tries=np.int32(len(fine)), **{"out_" + k: np.asarray(v, dtype=np.float32) for k, v in out.items()})


if __name__ == "__main__":
    main()
