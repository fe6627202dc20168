"""Evaluation-pass image metrics (utils/eval_util.py:28-49) and the resume-time parameter re-initialisation
(optimize_sequence.py:366-389)."""
import numpy as np
import pytest
import torch

import harp_oracle as O
from util import load_golden


def test_oracle_metrics_match_reference_golden():
    g = load_golden("eval_metrics")
    T = lambda k: torch.tensor(g[k])
    assert abs(float(O.sil_iou(T("ref_mask"), T("pred_mask"))) - float(g["sil_iou"])) < 1e-7
    assert abs(float(O.l1_diff(T("ref_image"), T("ref_mask"), T("pred_image"), T("pred_mask"))) - float(g["l1_diff"])) < 1e-7
    assert torch.equal(O.fill_bg(T("ref_image"), T("ref_mask")), T("fill_bg"))


def test_resume_params_host_logic_matches_oracle():
    """CPU tensors: the product's vectorised re-interpolation equals the reference's double loop bit for bit."""
    from harp_b200.file_utils import prepare_resumed_params
    gen = torch.Generator().manual_seed(2)
    p = dict(pose=torch.randn(97, 45, generator=gen), trans=torch.randn(97, 3, generator=gen), rot=torch.randn(97, 3, generator=gen),
             shape=torch.randn(10, generator=gen), texture=torch.rand(1, 8, 8, 3, generator=gen))
    mine, ref = prepare_resumed_params(p), O.resume_params(p)
    assert set(ref) <= set(mine)
    for k, v in ref.items():
        assert torch.equal(mine[k], v), k
    assert mine["wrist_pose"].shape == (97, 3) and float(mine["amb_ratio"]) == pytest.approx(0.4)
    assert mine["normal_map"].shape == (1, 512, 512, 3)
    assert not torch.equal(mine["pose"][:60], p["pose"][:60]) and torch.equal(mine["pose"][60:], p["pose"][60:])


@pytest.mark.gpu
def test_device_metrics_match_reference_golden():
    from harp_b200 import eval_util
    g = load_golden("eval_metrics")
    C = lambda k: torch.tensor(g[k]).cuda()
    assert abs(float(eval_util.sil_iou(C("ref_mask"), C("pred_mask"))) - float(g["sil_iou"])) < 1e-6
    assert abs(float(eval_util.l1_diff(C("ref_image"), C("ref_mask"), C("pred_image"), C("pred_mask"))) - float(g["l1_diff"])) < 1e-6
    assert torch.equal(eval_util.fill_bg(C("ref_image"), C("ref_mask")).cpu(), torch.tensor(g["fill_bg"]))
    st = eval_util.image_eval(dict(ref_image=[C("ref_image")[:2], C("ref_image")[2:]], pred_image=[C("pred_image")[:2], C("pred_image")[2:]],
                                   ref_mask=[C("ref_mask")[:2], C("ref_mask")[2:]], pred_mask=[C("pred_mask")[:2], C("pred_mask")[2:]]))
    assert abs(float(st["Silhouette IoU"]) - float(g["sil_iou"])) < 1e-6 and abs(float(st["L1"]) - float(g["l1_diff"])) < 1e-6
    # full size: 64 frames at 512^2 against torch on the device
    gen = torch.Generator(device="cuda").manual_seed(0)
    a, b = torch.rand(64, 512, 512, 3, device="cuda", generator=gen), torch.rand(64, 512, 512, 3, device="cuda", generator=gen)
    ma, mb = torch.rand(64, 512, 512, device="cuda", generator=gen), torch.rand(64, 512, 512, device="cuda", generator=gen)
    assert abs(float(eval_util.l1_diff(a, ma, b, mb)) - float((a - b).abs().double().mean())) < 1e-6
    assert abs(float(eval_util.sil_iou(ma, mb)) - float(O.sil_iou(ma.cpu(), mb.cpu()))) < 1e-6


@pytest.mark.gpu
def test_turntable_frames_match_oracle():
    """render_360 (utils/visualize.py:145-199): the rotated vertex sets and a sample of the 72 frames against the oracle."""
    from harp_b200 import synth, visualize, renderer_helper
    from harp_b200.manolayer import ManoLayer
    from util import golden_case
    g = load_golden("c2_full_128_sub")
    model, topo, params, batch, cfg = golden_case(g)
    tmpl = synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    layer = ManoLayer(pack, tmpl["faces_uvs"], tmpl["verts_uvs"])
    S, focal = 64, 1000.0 * 64 / 224.0
    light = params["light_position"][None]
    p = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in params.items()}
    p.update(faces_uvs=torch.tensor(tmpl["faces_uvs"]).long()[None].cuda(), verts_uvs=torch.tensor(tmpl["verts_uvs"])[None].cuda(),
             light_positions=light.cuda())
    fid = torch.arange(1)
    phong, _, _ = renderer_helper.get_renderers(image_size=S, light_posi=light.cuda(), device="cuda")
    configs = {"model_type": "harp", "img_size": S, "focal_length": focal, "use_arm": False}
    frames = visualize.render_360(p, fid, phong, S, focal, layer, configs, verts_textures=False, mesh_subdivider=object(), device="cuda")
    assert frames.shape == (72, S, S, 3) and frames.dtype == torch.uint8
    vm, _ = O.mano_forward(model, torch.cat((params["rot"][:1], params["pose"][:1]), 1), params["shape"][None], params["trans"][:1])
    verts = O.build_mesh(vm / 1000.0, topo, params["verts_disps"])
    tv = O.turntable_verts(verts)
    for i in (0, 17, 35, 36, 53, 71):
        ref = O.render_phong(tv[i], topo, params["cam"][:1], light, params["texture"], params["normal_map"], focal, S)
        ref8 = (ref[0, ..., :3].clip(0, 1) * 255).to(torch.uint8)
        d = (frames[i].cpu().int() - ref8.int()).abs().amax(-1)
        assert float((d <= 1).double().mean()) >= 0.99, (i, float((d <= 1).double().mean()))   # hard visibility at silhouette pixels


@pytest.mark.gpu
def test_light_sweep_frames_match_oracle():
    """render_360_light (utils/visualize.py:199-228): 40 Phong renders, no normal map, the light moving along z."""
    from harp_b200 import synth, visualize
    from harp_b200.manolayer import ManoLayer
    from harp_b200.structures import TexturesUV
    from util import golden_case
    g = load_golden("c2_full_128_sub")
    model, topo, params, batch, cfg = golden_case(g)
    tmpl = synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    layer = ManoLayer(pack, tmpl["faces_uvs"], tmpl["verts_uvs"])
    S, focal = 64, 1000.0 * 64 / 224.0
    vm, _ = O.mano_forward(model, torch.cat((params["rot"][:1], params["pose"][:1]), 1), params["shape"][None], params["trans"][:1])
    verts = O.build_mesh(vm / 1000.0, topo, params["verts_disps"])
    tex = TexturesUV(maps=params["texture"][:1].cuda(), faces_uvs=torch.tensor(tmpl["faces_uvs"]).long()[None].cuda(),
                     verts_uvs=torch.tensor(tmpl["verts_uvs"])[None].cuda())
    p = {"cam": params["cam"].cuda()}
    frames = visualize.render_360_light(p, torch.arange(1), verts.cuda(), layer.hc.faces()[None], tex, S, focal, device="cuda")
    assert frames.shape == (40, S, S, 3) and frames.dtype == torch.uint8
    flat = torch.tensor([0.0, 0.0, 1.0]).view(1, 1, 1, 3)
    for i in (0, 13, 39):
        light = torch.tensor([[1.0, 1.0, -5.0 + 10.0 / 40 * i]])
        ref = O.render_phong(verts, topo, params["cam"][:1], light, params["texture"], flat, focal, S)
        ref8 = (ref[0, ..., :3].clip(0, 1) * 255).to(torch.uint8)
        d = (frames[i].cpu().int() - ref8.int()).abs().amax(-1)
        assert float((d <= 1).double().mean()) >= 0.99, (i, float((d <= 1).double().mean()))
