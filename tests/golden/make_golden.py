"""Generate golden vectors by running the REFERENCE'S OWN, UNMODIFIED Python for the hot path.

Build-container only (needs /root/reference; the GPU box has no copy).  The reference files
    manopth/manolayer.py, utils/visualize.py (prepare_mesh / prepare_materials / render_image / render_image_with_RT),
    renderer/renderer_helper.py, renderer/pbr_materials.py, loss/arap.py, loss/kps_loss.py, loss/texture_reg.py
are imported from /root/reference and executed on CPU on top of the restated PyTorch3D shim (`oracle/p3d_shim`),
transcribing the loop body `optimize_sequence.py:446-569` call by call.  Outputs (images, loss scalars, parameter
gradients via autograd) are written to tests/golden/*.npz together with the inputs that produced them.

    python tests/golden/make_golden.py

What this pins: the oracle's L2/L3 restatement (`oracle/harp_oracle.py`) and, through it, the CUDA path.
What it cannot pin: the L1 PyTorch3D primitives themselves (restated in the shim; pytorch3d is not installable here).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "p3d_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from harp_b200 import synth  # noqa: E402

# torch-version drift, not a reference modification: `renderer_helper.py:397-399` clamps a LongTensor with
# `image_size-1` where image_size is a 0-dim *float* tensor (`visualize.py:312`).  Under the reference's pinned
# torch==1.11.0 (`requirements.txt:81`) the result stays integral and is used as an index; torch>=1.12 type-promotes
# clamp() to float and the indexing raises.  Emulate the pinned behaviour for integer tensors only.
_orig_clamp = torch.Tensor.clamp


def _clamp_compat(self, min=None, max=None):
    if not self.is_floating_point():
        if torch.is_tensor(min) and min.dim() == 0:
            min = int(min.item())
        if torch.is_tensor(max) and max.dim() == 0:
            max = int(max.item())
    return _orig_clamp(self, min, max)


torch.Tensor.clamp = _clamp_compat


def ref_mano_layer(pack):
    """ManoLayer without the licensed pkl: skip __init__'s load (manolayer.py:65), register the same buffers (70-97)."""
    from manopth.manolayer import ManoLayer
    layer = ManoLayer.__new__(ManoLayer)
    torch.nn.Module.__init__(layer)
    layer.center_idx, layer.robust_rot, layer.rot = None, False, 3
    layer.flat_hand_mean, layer.side, layer.use_pca = False, "right", False
    layer.joint_rot_mode, layer.root_rot_mode, layer.ncomps = "axisang", "axisang", 45
    for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights",
              "th_hands_mean"):
        layer.register_buffer(k, torch.tensor(pack[k]))
    layer.register_buffer("th_faces", torch.tensor(pack["th_faces"]).long())
    layer.register_buffer("th_selected_comps", torch.eye(45))
    layer.kintree_parents = synth.MANO_PARENTS
    return layer


_ARM_LAYER = {}


def ref_arm_layer(pack):
    """The reference's SMPLXARM (hand_models_harp/body_models.py) on the restated smplx modules, fed a synthetic
    SMPLX_NEUTRAL.npz; built as utils/hand_model_utils.py:65-70 does.  cwd must be /root/reference (body_models.py:2101)."""
    if "layer" not in _ARM_LAYER:
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import smplx_shim
        bm = smplx_shim.load_reference_body_models(REF)
        d = tempfile.mkdtemp()
        os.makedirs(os.path.join(d, "smplx"))
        np.savez(os.path.join(d, "smplx", "SMPLX_NEUTRAL.npz"), **pack)
        _ARM_LAYER["layer"] = bm.create(d, model_type="smplxarm", gender="neutral", use_face_contour=False, num_betas=10,
                                        use_pca=False, num_expression_coeffs=10, ext="npz")
    return _ARM_LAYER["layer"]


def reference_step(pack, tmpl, p, batch, S, focal, subdivide, stage, seed_albedo, seed_normal, arm=False, eval_renders=False):
    """optimize_sequence.py:446-569 with the reference's functions, device='cpu'."""
    import renderer.renderer_helper as renderer_helper
    from utils.visualize import prepare_mesh, prepare_materials, render_image, render_image_with_RT
    from loss.arap import arap_loss
    from loss.kps_loss import kps_loss
    from loss.texture_reg import albedo_reg, normal_reg
    from pytorch3d.structures import Meshes
    from pytorch3d.ops import SubdivideMeshes
    from pytorch3d.loss import mesh_laplacian_smoothing, mesh_normal_consistency

    dev = "cpu"
    layer = ref_arm_layer(pack) if arm else ref_mano_layer(pack)
    configs = {"model_type": "harp", "img_size": S, "focal_length": focal, "use_arm": arm,
               "share_light_position": True, "self_shadow": True, "known_appearance": False}
    mesh_subdivider = None
    if subdivide and arm:  # optimize_sequence.py:67-72
        hv, _ = layer(betas=torch.zeros([1, 10]), global_orient=torch.zeros([1, 3]), transl=torch.zeros([1, 3]),
                      right_hand_pose=torch.zeros([1, 45]), return_type="mano_w_arm")
        mesh_subdivider = SubdivideMeshes(meshes=Meshes(hv, layer.right_arm_faces_tensor.repeat(1, 1, 1)))
    elif subdivide:  # optimize_sequence.py:73-81
        hv, _ = layer(torch.zeros([1, 48]), torch.zeros([1, 10]), torch.zeros([1, 3]))
        mesh_subdivider = SubdivideMeshes(meshes=Meshes(hv, layer.th_faces.repeat(1, 1, 1)))
    params = dict(p)
    params["mesh_faces"] = layer.right_arm_faces_tensor if arm else layer.th_faces
    params["verts_uvs"] = torch.tensor(tmpl["verts_uvs"])[None]
    params["faces_uvs"] = torch.tensor(tmpl["faces_uvs"]).long()[None]
    B = p["pose"].shape[0]
    fid = torch.arange(B)
    l1_loss = torch.nn.L1Loss()
    light_positions = params["light_positions"][0].repeat(B, 1)
    phong_renderer, silhouette_renderer, normal_renderer = renderer_helper.get_renderers(
        image_size=S, light_posi=light_positions, silh_sigma=1e-7, silh_gamma=1e-1, silh_faces_per_pixel=50, device=dev)
    hand_joints, hand_verts, faces, textures = prepare_mesh(params, fid, layer, False, mesh_subdivider,
                                                            global_pose=False, configs=configs, shared_texture=True,
                                                            device=dev, use_arm=arm)
    meshes = Meshes(hand_verts, faces, textures)
    cam = params["cam"][fid]
    materials_properties = prepare_materials(params, B, device=dev)
    y_sil_pred = render_image(meshes, cam, B, silhouette_renderer, S, focal, silhouette=True, device=dev)
    from pytorch3d.renderer.mesh.rasterizer import MeshRasterizer
    sil_overflow = MeshRasterizer.last_overflow
    light_R, light_T, cam_R, cam_T = renderer_helper.process_info_for_shadow(
        cam, light_positions, hand_verts.mean(1), image_size=S, focal_length=focal, device=dev)
    shadow_renderer = renderer_helper.get_shadow_renderers(
        image_size=S, light_posi=light_positions, silh_sigma=1e-7, silh_gamma=1e-1, silh_faces_per_pixel=50,
        amb_ratio=torch.nn.Sigmoid()(params["amb_ratio"]), device=dev)
    y_pred = render_image_with_RT(meshes, light_T, light_R, cam_T, cam_R, B, shadow_renderer, S, focal,
                                  silhouette=False, materials_properties=materials_properties, device=dev)
    extra = {}
    if eval_renders:  # the two evaluation-pass renderers of get_renderers (optimize_sequence.py:486-488; visualize.py:145-228)
        with torch.no_grad():
            extra["phong"] = render_image(meshes, cam, B, phong_renderer, S, focal, silhouette=False, device=dev,
                                          materials_properties=materials_properties)
            extra["normals"] = render_image(meshes, cam, B, normal_renderer, S, focal, silhouette=False, device=dev,
                                            materials_properties=materials_properties)
    loss = {}
    if stage in ("coarse", "both"):
        loss["silhouette"] = l1_loss(batch["y_sil"], y_sil_pred)
        loss["kps_anchor"] = kps_loss(batch["init_joints"], hand_joints, use_arm=arm, device=dev)
        loss["vert_disp_reg"] = torch.sum(params["verts_disps"] ** 2.0)
        loss["laplacian"] = mesh_laplacian_smoothing(meshes)
        loss["normal"] = mesh_normal_consistency(meshes)
        loss["arap"] = arap_loss(meshes, Meshes(batch["ref_verts"], faces[:1]))
    if stage in ("both", "app"):
        me = batch["y_sil_eroded"].unsqueeze(-1)
        loss["photo"] = l1_loss(batch["y_true"] * me, y_pred * me)
        torch.manual_seed(seed_albedo)
        loss["albedo"] = albedo_reg(params["texture"], uv_mask=batch["uv_mask"], std=1.0)
        torch.manual_seed(seed_normal)
        loss["normal_reg"] = normal_reg(params["normal_map"], uv_mask=batch["uv_mask"])
    return loss, dict(sil=y_sil_pred, rgb=y_pred, verts=hand_verts, joints=hand_joints, sil_overflow=sil_overflow, **extra)


def scenario(name, S, B, subdivide, stage, dtype=torch.float32, trans_amp=0.0, arm=False, eval_renders=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import harp_oracle as O
    tmpl = synth.load_template("arm" if arm else "hand")
    pack = synth.make_smplx_pack(seed=0, template=tmpl) if arm else synth.make_mano_pack(seed=0, template=tmpl)
    focal = 1000.0 * S / 224.0
    seq = synth.make_sequence(max(B, 2), seed=1, img_size=S)
    app = synth.make_appearance(seed=2)
    V = (4083 if subdivide else 1026) if arm else (3093 if subdivide else 778)
    g = torch.Generator().manual_seed(7)
    trans = seq["trans"][:B] + trans_amp * torch.randn(B, 3, generator=g)
    P = dict(pose=seq["pose"][:B], rot=seq["rot"][:B], trans=trans, cam=seq["cam"][:B],
             shape=seq["shape"], verts_disps=0.0008 * torch.randn(V, 1, generator=g),
             texture=app["texture"], normal_map=app["normal_map"],
             light_positions=app["light_position"][None].repeat(B, 1), amb_ratio=app["amb_ratio"])
    if arm:
        # the arm template lies along -x with the wrist at the origin after centring (body_models.py:2342): a smaller
        # weak-perspective scale keeps the ~0.45 m arm inside the frame
        P["wrist_pose"] = 0.1 * torch.randn(B, 3, generator=g)
        P["cam"] = P["cam"] * torch.tensor([0.35, 1.0, 1.0]) + torch.tensor([0.0, 0.15, 0.0])
    uv_mask = torch.tensor(tmpl["uv_mask"])
    leaf = {k: torch.nn.Parameter(v.clone().to(dtype)) for k, v in P.items()}
    # targets: render a perturbed parameter set with the reference path, no grad
    with torch.no_grad():
        Pp = {k: v.clone() for k, v in P.items()}
        Pp["pose"] = Pp["pose"] + 0.05 * torch.randn(Pp["pose"].shape, generator=g)
        Pp["cam"] = Pp["cam"] + torch.tensor([0.1, 0.003, -0.002])
        app2 = synth.make_appearance(seed=5)
        Pp["texture"], Pp["normal_map"] = app2["texture"], app2["normal_map"]
        dummy = dict(y_sil=torch.zeros(B, S, S), y_sil_eroded=torch.zeros(B, S, S), y_true=torch.zeros(B, S, S, 3),
                     init_joints=torch.zeros(B, 21, 3), ref_verts=torch.zeros(1, V, 3), uv_mask=uv_mask)
        _, im = reference_step(pack, tmpl, Pp, dummy, S, focal, subdivide, "both", 11, 12, arm=arm)
        y_sil = (im["sil"] >= 0.5).float()
        y_er = synth.erode_mask(y_sil)
        y_true = im["rgb"].clamp(0, 1)
        init_joints = im["joints"][:, :21] * 1000.0
        d1 = {k: (v[:1] if k != "uv_mask" else v) for k, v in dummy.items()}
        _, im0 = reference_step(pack, tmpl, {k: v[:1] if k in ("pose", "rot", "trans", "cam", "light_positions", "wrist_pose") else v
                                             for k, v in P.items()}, d1, S, focal, subdivide, "both", 11, 12, arm=arm)
        ref_verts = im0["verts"][:1].clone()
    batch = dict(y_sil=y_sil, y_sil_eroded=y_er, y_true=y_true, init_joints=init_joints, ref_verts=ref_verts, uv_mask=uv_mask)
    loss, im = reference_step(pack, tmpl, leaf, batch, S, focal, subdivide, stage, 11, 12, arm=arm,
                              eval_renders=(name == "c2_full_128_sub") if eval_renders is None else eval_renders)
    total = sum(v * O.LOSS_WEIGHTS[k] for k, v in loss.items())
    names = list(leaf.keys())
    gs = torch.autograd.grad(total, [leaf[k] for k in names], allow_unused=True)
    out = {"S": S, "B": B, "subdivide": int(subdivide), "stage": stage, "focal": focal, "arm": int(arm),
           "seed_albedo": 11, "seed_normal": 12}
    for k, v in P.items():
        if k in ("texture", "normal_map"):  # regenerated from synth.make_appearance(seed=2); keep a checksum only
            out["insum_" + k] = np.float64(v.double().sum().item())
        else:
            out["in_" + k] = v.numpy()
    for k, v in batch.items():
        out["batch_" + k] = v.numpy()
    for k, v in loss.items():
        out["loss_" + k] = v.detach().numpy()
    out["loss_total"] = total.detach().numpy()
    out["img_sil"] = im["sil"].detach().numpy()
    out["img_rgb"] = im["rgb"].detach().numpy()
    out["verts"] = im["verts"].detach().numpy()
    out["joints"] = im["joints"].detach().numpy()
    out["sil_overflow"] = int(im["sil_overflow"])
    for k in ("phong", "normals"):
        if k in im:
            out["img_" + k] = im[k].detach().numpy()
    for k, g_ in zip(names, gs):
        out["grad_" + k] = (g_ if g_ is not None else torch.zeros_like(leaf[k])).detach().numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, {k: float(v) for k, v in loss.items()}, "overflow", im["sil_overflow"])


SCENARIOS = {
    "c1_sil_256_coarse": dict(S=256, B=1, subdivide=False, stage="coarse"),
    "c2_full_128_sub": dict(S=128, B=2, subdivide=True, stage="both", trans_amp=0.004),
    "c4_arm_128_sub": dict(S=128, B=2, subdivide=True, stage="both", trans_amp=0.004, arm=True),
    # BASELINE.json sizes and the reference's default operating point (utils/config_utils.py:7,11: use_arm, 448 x 448)
    "c2_full_512_sub": dict(S=512, B=2, subdivide=True, stage="both", trans_amp=0.004),
    "c4_arm_1024_sub": dict(S=1024, B=1, subdivide=True, stage="both", trans_amp=0.004, arm=True),
    "arm_448_sub": dict(S=448, B=2, subdivide=True, stage="both", trans_amp=0.004, arm=True),
}

if __name__ == "__main__":
    torch.set_num_threads(8)
    os.chdir(REF)
    for name in (sys.argv[1:] or list(SCENARIOS)):
        scenario(name, **SCENARIOS[name])
