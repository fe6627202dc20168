"""Shared helpers for the parity tests: rebuild oracle inputs from a golden .npz, error metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def golden_case(g, dtype=torch.float32):
    """-> (model, topo, params, batch, cfg) for oracle.step / the CUDA path."""
    import harp_oracle as O
    from harp_b200 import synth
    arm = bool(int(g["arm"])) if "arm" in g else False
    tmpl = synth.load_template("arm" if arm else "hand")
    pack = synth.make_smplx_pack(seed=0, template=tmpl) if arm else synth.make_mano_pack(seed=0, template=tmpl)
    model = {k: torch.tensor(v) for k, v in pack.items()}
    model = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in model.items()}
    sub = bool(int(g["subdivide"]))
    if arm:
        model["arm_vert"] = torch.tensor(tmpl["arm_vert"]).long()
        model["mano_joint"] = torch.tensor(tmpl["mano_joint"]).long()
        ids = synth.SMPLX_VERTEX_IDS  # VertexJointSelector order: face, feet, left tips, right tips
        names = (["nose", "reye", "leye", "rear", "lear", "LBigToe", "LSmallToe", "LHeel", "RBigToe", "RSmallToe", "RHeel"] +
                 [h + n for h in "lr" for n in ("thumb", "index", "middle", "ring", "pinky")])
        model["extra_joints_idxs"] = torch.tensor([ids[n] for n in names]).long()
    topo = O.Topology(tmpl["faces0"] if arm else pack["th_faces"], tmpl["faces_uvs"], tmpl["verts_uvs"], subdivide=sub)
    T = lambda a: torch.tensor(a).to(dtype) if np.issubdtype(a.dtype, np.floating) else torch.tensor(a)
    app = synth.make_appearance(seed=2)
    for k in ("texture", "normal_map"):
        assert abs(float(app[k].double().sum()) - float(g["insum_" + k])) < 1e-6, "synthetic appearance drifted"
    params = dict(pose=T(g["in_pose"]), rot=T(g["in_rot"]), trans=T(g["in_trans"]), cam=T(g["in_cam"]),
                  shape=T(g["in_shape"]), verts_disps=T(g["in_verts_disps"]), texture=app["texture"].to(dtype),
                  normal_map=app["normal_map"].to(dtype), light_position=T(g["in_light_positions"])[0],
                  amb_ratio=T(g["in_amb_ratio"]))
    if arm:
        params["wrist_pose"] = T(g["in_wrist_pose"])
    gen = lambda seed: (torch.manual_seed(int(seed)), torch.normal(mean=0, std=1.0, size=(512, 512, 2)).to(torch.int))[1]
    off_a = gen(g["seed_albedo"])
    # normal_reg draws after the same manual_seed; std default 2.0 (loss/texture_reg.py:33,48)
    torch.manual_seed(int(g["seed_normal"]))
    off_n = torch.normal(mean=0, std=2.0, size=(512, 512, 2)).to(torch.int)
    batch = dict(y_true=T(g["batch_y_true"]), y_sil=T(g["batch_y_sil"]), y_sil_eroded=T(g["batch_y_sil_eroded"]),
                 init_joints=T(g["batch_init_joints"]), ref_verts=T(g["batch_ref_verts"]), uv_mask=T(g["batch_uv_mask"]),
                 off_albedo=off_a, off_normal=off_n)
    cfg = dict(img_size=int(g["S"]), focal_length=float(g["focal"]), stage=str(g["stage"]), self_shadow=True, use_arm=arm)
    return model, topo, params, batch, cfg


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def frac_within(a, b, tol):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return float(((a - b).abs() <= tol).double().mean())


# ------------------------------------------------------------------------------------------------- measured-error log
# Every GPU parity test records the errors it MEASURED (not only the bound it asserts); tests/conftest.py writes them to
# $HARP_PARITY_OUT (default gpurun_out/parity.json) at session end, and the round's copy is committed under profiles/.
PARITY = {}


def record(test, **vals):
    d = PARITY.setdefault(test, {})
    for k, v in vals.items():
        d[k] = float(v) if isinstance(v, (int, float)) or (torch.is_tensor(v) and v.numel() == 1) else v


def agree_mask(a, b, tol=1e-4):
    """Pixels on which two renders of the same scene agree (per-pixel max over channels): the complement is the set of
    *flip* pixels, where a 1-ulp difference upstream changed a discrete decision (nearest face, band membership, rounded
    shadow tap) — see DESIGN.md section 2."""
    d = (torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs()
    if d.dim() == 4:
        d = d.amax(-1)
    return d <= tol
