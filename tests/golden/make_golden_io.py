"""Fixtures for the two file-facing "next" rows, produced by the REFERENCE'S OWN, unmodified functions.
Build-container only (needs /root/reference).

  f1  utils/data_util.py:load_img (11-30) on real PNG files written here -> tests/golden/io/*.png + load_img.npz
      (colour /255, mask /255, mask eroded by cv2.erode(3x3, iterations=2))
  f3  utils/file_utils.py:save_result / load_result (6-26) -> tests/golden/io/ref/saved_params.pkl written by the reference;
      this package's harp_b200/file_utils.py must read it, and the reference must read what this package writes (checked here
      live, and by tests/test_io_fixtures.py whenever /root/reference is present).

    python tests/golden/make_golden_io.py
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
OUT = os.path.join(HERE, "io")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


def params_dict(seed=0, T=3, V=12, H=8):
    """The keys of optimize_sequence.py:init_params (181-249), tiny shapes."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(trans=r(T, 3), pose=r(T, 45), rot=r(T, 3), wrist_pose=r(T, 3), shape=r(10), cam=r(T, 3), verts_disps=r(V, 1),
                verts_rgb=torch.rand(1, V, 3, generator=g), texture=torch.rand(1, H, H, 3, generator=g), normal_map=r(1, H, H, 3),
                light_positions=r(T, 3), amb_ratio=torch.tensor(0.4), mesh_faces=torch.randint(0, V, (7, 3), generator=g),
                verts_uvs=torch.rand(1, V + 3, 2, generator=g), faces_uvs=torch.randint(0, V + 3, (1, 7, 3), generator=g),
                uv_mask=(torch.rand(H, H, generator=g) > 0.5).float(), init_joints=r(T, 21, 3))


def main():
    os.makedirs(OUT, exist_ok=True)
    from utils.data_util import load_img            # the reference's own functions
    from utils import file_utils as ref_fu
    from harp_b200 import file_utils as our_fu
    rng = np.random.RandomState(5)
    S, out = 48, {}
    for i in range(2):
        rgb = rng.randint(0, 256, size=(S, S, 3)).astype(np.uint8)
        mask = (rng.rand(S, S) > 0.35).astype(np.uint8) * 255
        mask[8:20, 10:30] = 255
        mask[30:40, 5:25] = rng.randint(0, 256, size=(10, 20))  # grey values survive load_img un-thresholded
        ip, mp = os.path.join(OUT, f"frame_{i}.png"), os.path.join(OUT, f"mask_{i}.png")
        Image.fromarray(rgb).save(ip)
        Image.fromarray(mask, mode="L").save(mp)
        out[f"col_{i}"] = load_img(ip, torch_tensor=True).numpy()
        out[f"mask_{i}"] = load_img(mp, torch_tensor=True, load_mask=True).numpy()
        out[f"eroded_{i}"] = load_img(mp, torch_tensor=True, load_mask=True, erode=True).numpy()
    np.savez_compressed(os.path.join(OUT, "load_img.npz"), **out)
    # ---- checkpoints, both directions
    P = params_dict()
    os.makedirs(os.path.join(OUT, "ref"), exist_ok=True)
    ref_fu.save_result(P, os.path.join(OUT, "ref"))                       # written by the reference
    mine = our_fu.load_result(os.path.join(OUT, "ref"))                   # read by this package
    assert set(mine) == set(P) and all(torch.equal(mine[k], P[k]) for k in P), "harp_b200 cannot read the reference's checkpoint"
    tmp = os.path.join(OUT, "_ours")
    our_fu.save_result(P, tmp)                                            # written by this package
    # load_result calls set_require_grad(params) whose own default device is 'cuda' (file_utils.py:25,28): no GPU in the build
    # container, so the DEFAULT ARGUMENT is switched to 'cpu' for this call — the function body runs unmodified
    ref_fu.set_require_grad.__defaults__ = ("cpu",)
    back = ref_fu.load_result(tmp, device="cpu")                          # read by the reference
    assert set(back) == set(P) and all(torch.equal(back[k].detach(), P[k]) for k in P), "the reference cannot read harp_b200's checkpoint"
    assert isinstance(back["pose"], torch.nn.Parameter) and back["pose"].requires_grad
    import shutil
    shutil.rmtree(tmp)
    # ---- per-frame fit files (%04d_mano.pkl), both directions: hand_utils.write_pkl / load_params are the reference's own
    sys.path.insert(0, os.path.join(REF, "metro_modifications"))
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "trimesh", "imageio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    import hand_utils
    from harp_b200 import data_util as our_du
    g = torch.Generator().manual_seed(9)
    T = 4
    fit = dict(joints=torch.randn(T, 21, 3, generator=g), verts=torch.randn(T, 30, 3, generator=g), rot=torch.randn(T, 3, generator=g),
               pose=torch.randn(T, 45, generator=g), shape=torch.randn(T, 10, generator=g), trans=torch.randn(T, 3, generator=g),
               cam=torch.randn(T, 3, generator=g))
    d_ref = os.path.join(OUT, "mano_ref")
    os.makedirs(d_ref, exist_ok=True)
    hand_utils.write_pkl(fit, d_ref)                                     # written by the reference
    got = our_du.load_params(d_ref)                                       # read by this package
    assert all(torch.equal(got[k], fit[k]) for k in fit), "harp_b200 cannot read the reference's %04d_mano.pkl files"
    d_our = os.path.join(OUT, "_mano_ours")
    our_du.write_pkl(fit, d_our)                                          # written by this package
    back = hand_utils.load_params(d_our)                                  # read by the reference
    assert all(torch.equal(back[k], fit[k]) for k in fit), "the reference cannot read harp_b200's %04d_mano.pkl files"
    assert sorted(os.listdir(d_our)) == sorted(os.listdir(d_ref)) == ["%04d_mano.pkl" % i for i in range(1, T + 1)]
    spiky = dict(pose=fit["pose"].clone())
    spiky["pose"][2] += 3.0
    assert torch.equal(our_du.remove_spike(dict(pose=spiky["pose"].clone()))["pose"], hand_utils.remove_spike(dict(pose=spiky["pose"].clone()))["pose"])
    shutil.rmtree(d_our)
    print("load_img fixtures and checkpoint cross-read OK ->", OUT)


if __name__ == "__main__":
    main()
