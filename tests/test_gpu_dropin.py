"""GPU: the reference's loop body (optimize_sequence.py:446-569) written against harp_b200's drop-in modules — the same
call sequence tests/golden/make_golden.py runs against the reference's own files — compared with those goldens."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from util import load_golden, rel_l2, frac_within, record

_EXTRA = {}
LOSS_W = {"silhouette": 7.0, "kps_anchor": 10.0, "vert_disp_reg": 2.0, "normal": 0.1, "laplacian": 4.0, "arap": 0.2,
          "photo": 1.0, "albedo": 0.5, "normal_reg": 0.1}


def _loop_body(g, stage):
    from harp_b200 import synth
    from harp_b200 import renderer_helper
    from harp_b200.manolayer import ManoLayer
    from harp_b200.structures import Meshes
    from harp_b200.visualize import prepare_mesh, prepare_materials, render_image, render_image_with_RT
    from harp_b200.loss import arap_loss, kps_loss, albedo_reg, normal_reg, mesh_laplacian_smoothing, mesh_normal_consistency

    dev = "cuda"
    arm = bool(int(g["arm"])) if "arm" in g else False
    if arm:  # utils/hand_model_utils.py:63-70
        from harp_b200.body_models import SMPLXARM
        tmpl = synth.load_template("arm")
        layer = SMPLXARM(synth.make_smplx_pack(seed=0, template=tmpl), tmpl, device=dev)
    else:
        tmpl = synth.load_template("hand")
        pack = synth.make_mano_pack(seed=0, template=tmpl)
        layer = ManoLayer(pack, faces_uvs=tmpl["faces_uvs"], verts_uvs=tmpl["verts_uvs"], device=dev)
    S, focal, subdivide = int(g["S"]), float(g["focal"]), bool(int(g["subdivide"]))
    app = synth.make_appearance(seed=2)
    T = lambda a: torch.tensor(a).to(dev)
    keys = ("pose", "rot", "trans", "cam", "shape", "verts_disps", "light_positions", "amb_ratio") + (("wrist_pose",) if arm else ())
    leaf = {k: torch.nn.Parameter(T(g["in_" + k])) for k in keys}
    leaf["texture"] = torch.nn.Parameter(app["texture"].to(dev))
    leaf["normal_map"] = torch.nn.Parameter(app["normal_map"].to(dev))
    params = dict(leaf)
    params["mesh_faces"] = layer.right_arm_faces_tensor if arm else layer.th_faces
    params["verts_uvs"] = torch.tensor(tmpl["verts_uvs"])[None]
    params["faces_uvs"] = torch.tensor(tmpl["faces_uvs"]).long()[None]
    batch = {k[6:]: T(g[k]) for k in g if k.startswith("batch_")}
    B = leaf["pose"].shape[0]
    fid = torch.arange(B)
    configs = {"model_type": "harp", "img_size": S, "focal_length": focal, "use_arm": arm}
    mesh_subdivider = object() if subdivide else None
    l1_loss = torch.nn.L1Loss()
    # --- optimize_sequence.py:453-488 ---
    light_positions = params["light_positions"][0].repeat(B, 1)
    phong_renderer, silhouette_renderer, normal_renderer = renderer_helper.get_renderers(
        image_size=S, light_posi=light_positions, silh_sigma=1e-7, silh_gamma=1e-1, silh_faces_per_pixel=50, device=dev)
    hand_joints, hand_verts, faces, textures = prepare_mesh(params, fid, layer, False, mesh_subdivider, global_pose=False,
                                                            configs=configs, shared_texture=True, device=dev, use_arm=arm)
    meshes = Meshes(hand_verts, faces, textures)
    cam = params["cam"][fid]
    materials_properties = prepare_materials(params, B, device=dev)
    y_sil_pred = render_image(meshes, cam, B, silhouette_renderer, S, focal, silhouette=True, device=dev)
    light_R, light_T, cam_R, cam_T = renderer_helper.process_info_for_shadow(cam, light_positions, hand_verts.mean(1),
                                                                             image_size=S, focal_length=focal, device=dev)
    shadow_renderer = renderer_helper.get_shadow_renderers(image_size=S, light_posi=light_positions, silh_sigma=1e-7, silh_gamma=1e-1,
                                                           silh_faces_per_pixel=50, amb_ratio=torch.nn.Sigmoid()(params["amb_ratio"]),
                                                           device=dev)
    y_pred = render_image_with_RT(meshes, light_T, light_R, cam_T, cam_R, B, shadow_renderer, S, focal, silhouette=False,
                                  materials_properties=materials_properties, device=dev)
    # --- optimize_sequence.py:517-553 ---
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
        torch.manual_seed(int(g["seed_albedo"]))
        loss["albedo"] = albedo_reg(params["texture"], uv_mask=batch["uv_mask"], std=1.0)
        torch.manual_seed(int(g["seed_normal"]))
        loss["normal_reg"] = normal_reg(params["normal_map"], uv_mask=batch["uv_mask"])
    total = sum(v * LOSS_W[k] for k, v in loss.items())
    total.backward()
    _EXTRA.clear()
    if "img_phong" in g:  # the two evaluation-pass renderers, called exactly as the reference's eval code does
        with torch.no_grad():
            _EXTRA["phong"] = render_image(meshes, cam, B, phong_renderer, S, focal, silhouette=False, device=dev,
                                           materials_properties=materials_properties)
            _EXTRA["normals"] = render_image(meshes, cam, B, normal_renderer, S, focal, silhouette=False, device=dev,
                                             materials_properties=materials_properties)
    return loss, total, leaf, y_sil_pred, y_pred


@pytest.mark.parametrize("name", ["c2_full_128_sub", "c1_sil_256_coarse", "c4_arm_128_sub", "c2_full_512_sub", "arm_448_sub",
                                  "c4_arm_1024_sub"])
def test_dropin_loop_body_matches_reference_golden(name):
    g = load_golden(name)
    loss, total, leaf, sil, rgb = _loop_body(g, str(g["stage"]))
    for k, v in loss.items():
        assert math.isclose(float(v), float(g["loss_" + k]), rel_tol=2e-3, abs_tol=1e-7), (k, float(v), float(g["loss_" + k]))
    assert math.isclose(float(total), float(g["loss_total"]), rel_tol=1e-3)
    d = (sil.detach().cpu() - torch.tensor(g["img_sil"])).abs()
    assert float((d <= 1e-4).double().mean()) >= 0.998
    d = (rgb.detach().cpu() - torch.tensor(g["img_rgb"])).abs().amax(-1)
    assert float((d <= 1e-4).double().mean()) >= 0.99
    S_, B_ = int(g["S"]), int(g["B"])
    if float(g["grad_amb_ratio"]) != 0.0:  # cancelling scalar sum of per-channel sign quanta (see test_gpu_kernels.py)
        assert abs(float(leaf["amb_ratio"].grad) - float(g["grad_amb_ratio"])) <= 40 * 2.0 / (3 * B_ * S_ * S_) / 4
    for k in ("pose", "rot", "cam", "shape", "verts_disps", "light_positions", "trans", "wrist_pose"):
        if "grad_" + k not in g:
            continue
        gg = torch.tensor(g["grad_" + k])
        if float(gg.abs().max()) == 0.0:
            continue
        assert leaf[k].grad is not None, k
        r = rel_l2(leaf[k].grad.cpu().view(gg.shape), gg)
        record(f"dropin_vs_golden[{name}]", **{"grad_" + k: r})
        assert r < 2e-2, (k, r)
    for k in ("texture", "normal_map"):
        gg = torch.tensor(g["grad_" + k])
        if float(gg.abs().max()) == 0.0:
            continue
        assert frac_within(leaf[k].grad.cpu().view(gg.shape), gg, 1e-3 * float(gg.abs().max())) >= 0.9995, k
    for k, img in _EXTRA.items():  # SoftPhongShaderPBR and the K=10 normal renderer vs the reference's own renders
        d = (img.cpu() - torch.tensor(g["img_" + k])).abs().amax(-1)
        assert float((d <= 2e-4).double().mean()) >= 0.995, (k, float((d <= 2e-4).double().mean()))
        assert float((d <= 2e-2).double().mean()) >= 0.999, k


def test_alpha_channel_is_the_coverage_not_a_colour_test():
    """The K=1 renderers' alpha channel (renderer_helper.py:190,592) comes from the rasteriser's hit list: a hand shaded exactly
    white (white texture, ambient 1) on the white background keeps its alpha, and the mask equals the set of pixels the
    silhouette's hard interior covers."""
    from harp_b200 import camera, ops, scene, synth
    hc, pack, tmpl = scene.make_context("cuda:0")
    S, B = 128, 2
    focal = 1000.0 * S / 224.0
    seq = synth.make_sequence(B, seed=3, img_size=S)
    dev = hc.device
    pose48 = torch.cat((seq["rot"], seq["pose"]), 1).to(dev)
    verts0, _ = ops.LbsMano.apply(hc, pose48, seq["shape"].to(dev), seq["trans"].to(dev))
    verts, normals = ops.MeshBuild.apply(hc, verts0, torch.zeros(hc.V, 1, device=dev))
    cR, cT = camera.flip_xy_R(B, dev), camera.cam_to_T(seq["cam"].to(dev), focal, S)
    light = torch.tensor([[1.0, 1.0, -5.0]], device=dev).expand(B, 3).contiguous()
    lR, lT = camera.light_camera(light, verts.mean(1))
    app = synth.make_appearance(seed=2)
    masks = []
    for tex, amb in ((app["texture"].to(dev), 0.6), (torch.ones_like(app["texture"]).to(dev), 1.0)):
        rgb = ops.Shade.apply(hc, verts, normals, cR, cT, lR, lT, light, torch.tensor([amb], device=dev), tex,
                              app["normal_map"].to(dev), focal, S, 1)
        masks.append(ops.hit_mask(hc, B, S))
    assert float((rgb - 1.0).abs().max()) < 1e-6  # the second render is white on white (to the rounding of the bilinear weights) ...
    assert torch.equal(masks[0], masks[1]) and masks[1].sum() > 500  # ... and still has its alpha
    assert torch.equal(masks[0][..., 0] > 0, (ops.Shade.apply(hc, verts, normals, cR, cT, lR, lT, light, torch.tensor([0.6], device=dev),
                                                               app["texture"].to(dev), app["normal_map"].to(dev), focal, S, 1) != 1.0).any(-1))
    alpha = ops.Silhouette.apply(hc, verts, cR, cT, focal, S, 1e-7)
    inter = (masks[0][..., 0] > 0) & (alpha < 0.5)
    assert int(inter.sum()) == 0  # every covered pixel centre is inside the soft silhouette
