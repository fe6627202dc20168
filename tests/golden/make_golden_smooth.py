"""Golden vector for the upstream sequence smoothing: runs the REFERENCE'S OWN, UNMODIFIED
`metro_modifications/hand_utils.py:optimize_smooth_seq` (540-688) on CPU.  Build-container only.  Environment shims (none
touches the reference's arithmetic): placeholder modules for its plotting / mesh imports, `.to("cuda")` mapped to the CPU
(no GPU here), and `ReduceLROnPlateau(verbose=True)` (652) accepted again — torch >= 2.7 removed that keyword.

    python tests/golden/make_golden_smooth.py      # ~1 min of CPU
"""
import contextlib
import io
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
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

_t_to, _m_to = torch.Tensor.to, torch.nn.Module.to
_cpu = lambda args: tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
torch.Tensor.to = lambda self, *a, **k: _t_to(self, *_cpu(a), **k)
torch.nn.Module.to = lambda self, *a, **k: _m_to(self, *_cpu(a), **k)

_Plateau = torch.optim.lr_scheduler.ReduceLROnPlateau


class _PlateauCompat(_Plateau):
    def __init__(self, optimizer, *a, verbose=False, **k):
        super().__init__(optimizer, *a, **k)


torch.optim.lr_scheduler.ReduceLROnPlateau = _PlateauCompat

from harp_b200 import synth  # noqa: E402
from make_golden import ref_mano_layer, ref_arm_layer  # noqa: E402


def main():
    import hand_utils
    torch.set_num_threads(os.cpu_count() or 1)
    tmpl = synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    layer = ref_mano_layer(pack)
    T = 12
    seq = synth.make_sequence(T, seed=1)
    g = torch.Generator().manual_seed(5)
    # per-frame fits as the upstream stage leaves them: a smooth motion plus per-frame jitter
    rot = seq["rot"] + 0.03 * torch.randn(T, 3, generator=g)
    pose = seq["pose"] + 0.05 * torch.randn(T, 45, generator=g)
    shape = seq["shape"][None].repeat(T, 1) + 0.1 * torch.randn(T, 10, generator=g)
    trans = torch.tensor([[0.0, 0.0, 0.5]]).repeat(T, 1) + 0.002 * torch.randn(T, 3, generator=g)
    cam = seq["cam"] + torch.tensor([0.3, 0.01, 0.01]) * torch.randn(T, 3, generator=g)
    with torch.no_grad():
        verts, joints = layer(torch.cat((rot, pose), 1), shape, trans)
    params_in = dict(rot=rot, pose=pose, shape=shape, trans=trans, cam=cam, joints=joints, verts=verts)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = hand_utils.optimize_smooth_seq({k: v.clone() for k, v in params_in.items()}, layer, img_res=224)
    log = buf.getvalue()
    n_pose_iters = sum(1 for l in log.splitlines() if l.startswith("smooth "))
    print(log[-300:])
    np.savez_compressed(os.path.join(HERE, "smooth_seq_t12.npz"), **{"in_" + k: v.numpy() for k, v in params_in.items()},
                        **{"out_" + k: v.numpy() for k, v in out.items()}, printed_pose_lines=np.int32(n_pose_iters))
    print({k: tuple(v.shape) for k, v in out.items()})
    print("moved: pose", float((out["pose"] - pose).abs().max()), "cam", float((out["cam"] - cam).abs().max()))


def main_arm():
    """The reference's default branch (utils/config_utils.py:7 use_arm=True): optimize_smooth_seq(use_smplx_arm=True) on the
    reference's own SMPLXARM layer (hand_models_harp/body_models.py, return_type='mano'), synthetic SMPL-X-format model."""
    import hand_utils
    torch.set_num_threads(os.cpu_count() or 1)
    os.chdir(REF)  # body_models.py:2101 opens the arm correspondence relative to the reference root
    tmpl = synth.load_template("arm")
    pack = synth.make_smplx_pack(seed=0, template=tmpl)
    layer = ref_arm_layer(pack)
    T = 8
    seq = synth.make_sequence(T, seed=1)
    g = torch.Generator().manual_seed(6)
    rot = seq["rot"] + 0.03 * torch.randn(T, 3, generator=g)
    pose = seq["pose"] + 0.05 * torch.randn(T, 45, generator=g)
    shape = seq["shape"][None].repeat(T, 1) + 0.1 * torch.randn(T, 10, generator=g)
    trans = torch.tensor([[0.0, 0.0, 0.5]]).repeat(T, 1) + 0.002 * torch.randn(T, 3, generator=g)
    cam = seq["cam"] + torch.tensor([0.3, 0.01, 0.01]) * torch.randn(T, 3, generator=g)
    with torch.no_grad():
        verts, joints = layer(betas=shape, global_orient=rot, transl=trans, right_hand_pose=pose, return_type="mano")
    params_in = dict(rot=rot, pose=pose, shape=shape, trans=trans, cam=cam, joints=joints, verts=verts)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = hand_utils.optimize_smooth_seq({k: v.clone() for k, v in params_in.items()}, layer, use_smplx_arm=True, img_res=224,
                                             total_iter_pose=300)
    print(buf.getvalue()[-200:])
    np.savez_compressed(os.path.join(HERE, "smooth_seq_arm_t8.npz"), **{"in_" + k: v.numpy() for k, v in params_in.items()},
                        **{"out_" + k: v.numpy() for k, v in out.items()})
    print({k: tuple(v.shape) for k, v in out.items()})
    print("moved: pose", float((out["pose"] - pose).abs().max()), "cam", float((out["cam"] - cam).abs().max()))


if __name__ == "__main__":
    if "arm" in sys.argv[1:]:
        main_arm()
    else:
        main()
