"""GPU parity: every CUDA kernel pair, called through the C ABI (ctypes), against the CPU oracle on identical inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import harp_oracle as O
from util import load_golden, golden_case, rel_l2, frac_within, record


def _arm_ctx(sub=True):
    from harp_b200 import synth
    from harp_b200.ops import HarpContext
    tmpl = synth.load_template("arm")
    red = synth.reduce_smplx_arm(synth.make_smplx_pack(seed=0, template=tmpl), tmpl)
    return HarpContext(tmpl["faces0"], subdivide=sub, faces_uvs=tmpl["faces_uvs"], verts_uvs=tmpl["verts_uvs"], lbs=red)


def _ctx(model, topo_sub, tmpl=None, with_uv=True):
    from harp_b200 import synth
    from harp_b200.ops import HarpContext
    if "kintree_table" in model:
        return _arm_ctx(topo_sub)
    tmpl = tmpl or synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    return HarpContext(pack["th_faces"], subdivide=topo_sub, faces_uvs=tmpl["faces_uvs"] if with_uv else None,
                       verts_uvs=tmpl["verts_uvs"] if with_uv else None, mano=pack)


@pytest.fixture(scope="module")
def case():
    g = load_golden("c2_full_128_sub")
    return golden_case(g)


def test_lbs_forward_backward(case):
    from harp_b200.ops import LbsMano
    model, topo, params, batch, cfg = case
    hc = _ctx(model, True)
    B = params["pose"].shape[0]
    pose48 = torch.cat((params["rot"], params["pose"]), 1)
    p_o = pose48.clone().requires_grad_(True); b_o = params["shape"].clone().requires_grad_(True); t_o = params["trans"].clone().requires_grad_(True)
    v_o, j_o = O.mano_forward(model, p_o, b_o.repeat(B, 1), t_o)
    v_o, j_o = v_o / 1000, j_o / 1000
    gv, gj = torch.randn_like(v_o), torch.randn_like(j_o)
    (v_o * gv).sum().add((j_o * gj).sum()).backward()
    p_c = pose48.cuda().requires_grad_(True); b_c = params["shape"].cuda().requires_grad_(True); t_c = params["trans"].cuda().requires_grad_(True)
    v_c, j_c = LbsMano.apply(hc, p_c, b_c, t_c)
    ((v_c * gv.cuda()).sum() + (j_c * gj.cuda()).sum()).backward()
    record("lbs_mano", verts_maxabs_m=float((v_c.cpu() - v_o).abs().max()), joints_maxabs_m=float((j_c.cpu() - j_o).abs().max()),
           grad_pose=rel_l2(p_c.grad.cpu(), p_o.grad), grad_betas=rel_l2(b_c.grad.cpu(), b_o.grad), grad_trans=rel_l2(t_c.grad.cpu(), t_o.grad))
    assert (v_c.cpu() - v_o).abs().max() < 2e-6      # metres
    assert (j_c.cpu() - j_o).abs().max() < 2e-6
    assert rel_l2(p_c.grad.cpu(), p_o.grad) < 1e-4
    assert rel_l2(b_c.grad.cpu(), b_o.grad) < 1e-4
    assert rel_l2(t_c.grad.cpu(), t_o.grad) < 1e-4


def test_lbs_smplx_arm_forward_backward():
    """Generic LBS kernel on the SMPL-X arm (55 joints, 1026 kept verts, wrist-centred, 22 output joints) against the
    oracle's SMPLXARM.forward restatement, which the reference's own body_models.py reproduces bit for bit."""
    from harp_b200.body_models import SMPLXARM
    from harp_b200.ops import Lbs
    g = load_golden("c4_arm_128_sub")
    model, topo, params, batch, cfg = golden_case(g)
    hc = _arm_ctx(True)
    B = params["pose"].shape[0]
    names = ("shape", "rot", "pose", "wrist_pose", "trans")
    o = {k: params[k].clone().requires_grad_(True) for k in names}
    v_o, j_o = O.smplxarm_forward(model, o["shape"].repeat(B, 1), o["rot"], o["pose"], o["wrist_pose"], o["trans"])
    v_o, j_o = v_o / 1000, j_o / 1000
    torch.manual_seed(2)
    gv, gj = torch.randn_like(v_o), torch.randn_like(j_o)
    ((v_o * gv).sum() + (j_o * gj).sum()).backward()
    c = {k: params[k].clone().cuda().requires_grad_(True) for k in names}
    v_c, j_c = Lbs.apply(hc, SMPLXARM.assemble_full_pose(c["rot"], c["pose"], c["wrist_pose"]), c["shape"], c["trans"])
    ((v_c * gv.cuda()).sum() + (j_c * gj.cuda()).sum()).backward()
    record("lbs_smplx_arm", verts_maxabs_m=float((v_c.cpu() - v_o).abs().max()), joints_maxabs_m=float((j_c.cpu() - j_o).abs().max()),
           grads={k: rel_l2(c[k].grad.cpu(), o[k].grad) for k in names})
    assert (v_c.cpu() - v_o).abs().max() < 3e-6 and (j_c.cpu() - j_o).abs().max() < 3e-6
    for k in names:
        assert rel_l2(c[k].grad.cpu(), o[k].grad) < 2e-4, (k, rel_l2(c[k].grad.cpu(), o[k].grad))


@pytest.mark.parametrize("sub", [True, False])
def test_mesh_build_forward_backward(case, sub):
    from harp_b200.ops import MeshBuild
    import pytorch3d._core as p3d
    model, topo, params, batch, cfg = case
    hc = _ctx(model, sub)
    topo_o = O.Topology(model["th_faces"], subdivide=sub)
    torch.manual_seed(0)
    B = 3
    v0 = model["th_v_template"].repeat(B, 1, 1) + 0.002 * torch.randn(B, 778, 3)
    disps = 0.001 * torch.randn(topo_o.V, 1)
    v0_o, d_o = v0.clone().requires_grad_(True), disps.clone().requires_grad_(True)
    verts_o = O.build_mesh(v0_o, topo_o, d_o)
    n_o = p3d.Meshes(verts_o, topo_o.faces[None].expand(B, -1, -1)).verts_normals_padded()
    gv, gn = torch.randn_like(verts_o), torch.randn_like(n_o)
    ((verts_o * gv).sum() + (n_o * gn).sum()).backward()
    v0_c, d_c = v0.cuda().requires_grad_(True), disps.cuda().requires_grad_(True)
    verts_c, n_c = MeshBuild.apply(hc, v0_c, d_c)
    ((verts_c * gv.cuda()).sum() + (n_c * gn.cuda()).sum()).backward()
    record(f"mesh_build[sub={sub}]", verts_maxabs_m=float((verts_c.cpu() - verts_o).abs().max()),
           normals_maxabs=float((n_c.cpu() - n_o).abs().max()), grad_verts0=rel_l2(v0_c.grad.cpu(), v0_o.grad),
           grad_disps=rel_l2(d_c.grad.cpu(), d_o.grad))
    assert (hc.faces().cpu() == topo_o.faces).all()
    assert (verts_c.cpu() - verts_o).abs().max() < 1e-6
    assert (n_c.cpu() - n_o).abs().max() < 2e-4
    assert rel_l2(v0_c.grad.cpu(), v0_o.grad) < 2e-4
    assert rel_l2(d_c.grad.cpu(), d_o.grad) < 2e-4


def test_mesh_losses(case):
    from harp_b200.ops import MeshLosses
    model, topo, params, batch, cfg = case
    hc = _ctx(model, True)
    torch.manual_seed(1)
    B = 2
    verts = batch["ref_verts"].repeat(B, 1, 1) + 0.0005 * torch.randn(B, topo.V, 3)
    joints = 0.05 * torch.randn(B, 21, 3)
    init = 50 * torch.randn(B, 21, 3)
    disps = 0.001 * torch.randn(topo.V, 1)
    w = torch.tensor([4.0, 0.1, 0.2, 10.0, 2.0])
    v_o, j_o, d_o = verts.clone().requires_grad_(True), joints.clone().requires_grad_(True), disps.clone().requires_grad_(True)
    lap, nc = O.mesh_reg_losses(v_o, topo)
    lo = torch.stack([lap, nc, O.arap_loss(v_o, batch["ref_verts"], topo), O.kps_loss(init, j_o), torch.sum(d_o ** 2.0)])
    (lo * w).sum().backward()
    v_c, j_c, d_c = verts.cuda().requires_grad_(True), joints.cuda().requires_grad_(True), disps.cuda().requires_grad_(True)
    lc = MeshLosses.apply(hc, v_c, batch["ref_verts"].cuda(), j_c, init.cuda(), d_c)
    (lc * w.cuda()).sum().backward()
    record("mesh_losses", losses_rel=[abs(float(lc[i]) - float(lo[i])) / abs(float(lo[i])) for i in range(5)],
           grad_verts=rel_l2(v_c.grad.cpu(), v_o.grad), grad_joints=rel_l2(j_c.grad.cpu(), j_o.grad), grad_disps=rel_l2(d_c.grad.cpu(), d_o.grad))
    for i in range(5):
        assert math.isclose(float(lc[i]), float(lo[i]), rel_tol=2e-4), (i, float(lc[i]), float(lo[i]))
    assert rel_l2(v_c.grad.cpu(), v_o.grad) < 5e-4
    assert rel_l2(j_c.grad.cpu(), j_o.grad) < 1e-4
    assert rel_l2(d_c.grad.cpu(), d_o.grad) < 1e-5


@pytest.mark.parametrize("S,sub,K", [(128, True, 50), (256, False, 50), (64, True, 50), (128, True, 3), (64, True, 8)])
def test_silhouette_forward_backward(case, S, sub, K):
    """Same world-space verts into both paths.  alpha is hard 0/1 except in the +-1/3 px bands, where the sigma=1e-7
    sigmoid amplifies 1-ulp differences of the projected coordinates; both tolerances are stated.
    K < number of candidate faces on a pixel (sub-pixel faces at 64^2, or an artificially small K) exercises PyTorch3D's
    keep-the-K-nearest truncation, which the CUDA path reproduces with its exact per-pixel top-K kernel."""
    from harp_b200.ops import Silhouette
    model, topo, params, batch, cfg = case
    hc = _ctx(model, sub)
    topo_o = O.Topology(model["th_faces"], subdivide=sub)
    B = 2
    focal = 1000.0 * S / 224.0
    pose48 = torch.cat((params["rot"], params["pose"]), 1)
    vm, _ = O.mano_forward(model, pose48, params["shape"].repeat(B, 1), params["trans"])
    verts = O.build_mesh(vm / 1000.0, topo_o, None).detach()
    cam = params["cam"]
    v_o, cam_o = verts.clone().requires_grad_(True), cam.clone().requires_grad_(True)
    a_o, ov = O.render_silhouette(v_o, topo_o, cam_o, focal, S, K=K, return_overflow=True)
    torch.manual_seed(3)
    g = torch.randn_like(a_o)
    (a_o * g).sum().backward()
    v_c, cam_c = verts.cuda().requires_grad_(True), cam.cuda().requires_grad_(True)
    R = O.cam_R(B, torch.float32).cuda()
    T = O.cam_to_T(cam_c, focal, S)
    a_c = Silhouette.apply(hc, v_c, R, T, focal, S, 1e-7, K)
    (a_c * g.cuda()).sum().backward()
    cnt = hc.counters()
    assert cnt["bin_overflow"] == 0 and cnt["sil_k_overflow_pixels"] == 0
    if K < 50:
        assert ov, "this case is meant to exercise the top-K truncation"
    d = (a_c.cpu() - a_o.detach()).abs()
    record(f"silhouette[S={S},sub={sub},K={K}]", alpha_frac_1e4=float((d <= 1e-4).double().mean()),
           alpha_frac_5e3=float((d <= 5e-3).double().mean()), grad_verts=rel_l2(v_c.grad.cpu(), v_o.grad),
           grad_cam=rel_l2(cam_c.grad.cpu(), cam_o.grad))
    assert float((d <= 1e-4).double().mean()) >= 0.999, float((d <= 1e-4).double().mean())
    assert float((d <= 5e-3).double().mean()) >= 0.9995
    assert rel_l2(v_c.grad.cpu(), v_o.grad) < 2e-2
    assert rel_l2(cam_c.grad.cpu(), cam_o.grad) < 2e-2


@pytest.mark.parametrize("S,shadow", [(128, True), (64, True), (128, False)])
def test_shade_forward_backward(case, S, shadow):
    """Colour render: hard visibility (a 1-ulp vertex move can flip a pixel's face) => image tolerance on a pixel
    fraction; gradients are sums over ~1e4 pixels and must agree to 1e-3 in relative L2."""
    import pytorch3d._core as p3d
    from harp_b200.ops import Shade
    from harp_b200 import camera
    model, topo, params, batch, cfg = case
    hc = _ctx(model, True)
    B = 2
    focal = 1000.0 * S / 224.0
    pose48 = torch.cat((params["rot"], params["pose"]), 1)
    vm, _ = O.mano_forward(model, pose48, params["shape"].repeat(B, 1), params["trans"])
    verts = O.build_mesh(vm / 1000.0, topo, params["verts_disps"]).detach()
    names = ("verts", "cam", "light", "amb", "tex", "nm")
    base = dict(verts=verts, cam=params["cam"], light=params["light_position"][None].repeat(B, 1).clone(),
                amb=params["amb_ratio"], tex=params["texture"], nm=params["normal_map"])
    o = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    rgb_o = O.render_rgb_shadow(o["verts"], topo, o["cam"], o["light"], o["amb"], o["tex"], o["nm"], focal, S, shadow)
    torch.manual_seed(5)
    g = torch.randn_like(rgb_o)
    (rgb_o * g).sum().backward()
    c = {k: v.clone().cuda().requires_grad_(True) for k, v in base.items()}
    faces = topo.faces.cuda()[None].expand(B, -1, -1)
    normals = p3d.Meshes(c["verts"], faces).verts_normals_padded()
    cR = camera.flip_xy_R(B, "cuda")
    cT = camera.cam_to_T(c["cam"], focal, S)
    lR, lT = camera.light_camera(c["light"], c["verts"].mean(1))
    rgb_c = Shade.apply(hc, c["verts"], normals, cR, cT, lR, lT, c["light"], torch.sigmoid(c["amb"]), c["tex"], c["nm"],
                        focal, S, int(shadow))
    (rgb_c * g.cuda()).sum().backward()
    assert hc.counters()["bin_overflow"] == 0
    d = (rgb_c.detach().cpu() - rgb_o.detach()).abs().amax(-1)
    record(f"shade[S={S},shadow={shadow}]", rgb_frac_1e4=float((d <= 1e-4).double().mean()), rgb_frac_2e2=float((d <= 2e-2).double().mean()),
           grads={k: rel_l2(c[k].grad.cpu(), o[k].grad) for k in names})
    assert float((d <= 1e-4).double().mean()) >= 0.995, float((d <= 1e-4).double().mean())
    assert float((d <= 2e-2).double().mean()) >= 0.9995
    for k in names:
        r = rel_l2(c[k].grad.cpu(), o[k].grad)
        assert r < (5e-3 if k in ("verts",) else 2e-3), (k, r)


def test_phong_twin_and_normal_renderer(case):
    """Non-shadow SoftPhongShaderPBR (fixed lights .5/.4/.1, differentiable) and the K=10 normal renderer (forward)."""
    import pytorch3d._core as p3d
    from harp_b200.ops import Shade, render_normals
    from harp_b200 import camera
    model, topo, params, batch, cfg = case
    hc = _ctx(model, True)
    B, S = 2, 128
    focal = 1000.0 * S / 224.0
    pose48 = torch.cat((params["rot"], params["pose"]), 1)
    vm, _ = O.mano_forward(model, pose48, params["shape"].repeat(B, 1), params["trans"])
    verts = O.build_mesh(vm / 1000.0, topo, params["verts_disps"]).detach()
    light = params["light_position"][None].repeat(B, 1).clone()
    base = dict(verts=verts, cam=params["cam"], light=light, tex=params["texture"], nm=params["normal_map"])
    o = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    rgb_o = O.render_phong(o["verts"], topo, o["cam"], o["light"], o["tex"], o["nm"], focal, S)
    torch.manual_seed(6)
    g = torch.randn_like(rgb_o)
    (rgb_o * g).sum().backward()
    c = {k: v.clone().cuda().requires_grad_(True) for k, v in base.items()}
    faces = topo.faces.cuda()[None].expand(B, -1, -1)
    normals = p3d.Meshes(c["verts"], faces).verts_normals_padded()
    cR, cT = camera.flip_xy_R(B, "cuda"), camera.cam_to_T(c["cam"], focal, S)
    rgb_c = Shade.apply(hc, c["verts"], normals, cR, cT, cR, cT, c["light"], torch.zeros(1, device="cuda"), c["tex"], c["nm"],
                        focal, S, 0, (0.5, 0.4, 0.1))
    (rgb_c * g.cuda()).sum().backward()
    d = (rgb_c.detach().cpu() - rgb_o.detach()).abs().amax(-1)
    assert float((d <= 1e-4).double().mean()) >= 0.995
    for k in base:
        r = rel_l2(c[k].grad.cpu(), o[k].grad)
        assert r < 5e-3, (k, r)
    with torch.no_grad():
        n_o = O.render_normals(verts, topo, params["cam"], params["normal_map"], focal, S, K=10)
        n_c = render_normals(hc, verts.cuda(), normals.detach(), cR, cT.detach(), focal, S, params["normal_map"].cuda(), 10)
    d = (n_c.cpu() - n_o).abs().amax(-1)
    assert float((d <= 2e-4).double().mean()) >= 0.995 and float((d <= 2e-2).double().mean()) >= 0.999


def test_texture_reg_and_adam():
    from harp_b200 import synth
    from harp_b200.ops import TextureReg, adam_step
    app = synth.make_appearance(seed=2)
    mask = torch.tensor(synth.load_template("hand")["uv_mask"])
    off = synth.make_reg_offsets(seed=3, std=1.0)
    for kind, key in ((0, "texture"), (1, "normal_map")):
        t_o = app[key].clone().requires_grad_(True)
        lo = O.smooth_texture_reg(t_o, off, mask) if kind == 0 else O.normal_reg(t_o, off, mask)
        lo.backward()
        t_c = app[key].clone().cuda().requires_grad_(True)
        lc = TextureReg.apply(None, kind, t_c, off.cuda(), mask.cuda())
        lc.backward()
        assert math.isclose(float(lc), float(lo), rel_tol=1e-4)
        assert rel_l2(t_c.grad.cpu(), t_o.grad) < 1e-4
    torch.manual_seed(0)
    p = torch.randn(1000, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-2)
    pc, m, v = p.detach().clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for step in range(1, 6):
        gr = torch.randn(1000)
        p.grad = gr.clone()
        opt.step()
        adam_step(pc, gr.cuda(), m, v, 1e-2, step)
    assert (pc.cpu() - p.detach()).abs().max() < 1e-6


def test_adam_device_step_counter_matches_torch():
    from harp_b200 import _lib
    torch.manual_seed(0)
    p = torch.randn(777, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-3)
    pc, m, v = p.detach().clone().cuda(), torch.zeros(777).cuda(), torch.zeros(777).cuda()
    state = torch.zeros(3, dtype=torch.int32, device="cuda")
    for _ in range(4):
        gr = torch.randn(777)
        p.grad = gr.clone()
        opt.step()
        g = gr.cuda()
        _lib.check(_lib.lib().harp_adam_dev(pc.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 777, 1e-3, 0.9, 0.999, 1e-8,
                                            state.data_ptr(), _lib.stream()))
    assert int(state[0]) == 4
    assert (pc.cpu() - p.detach()).abs().max() < 1e-6


def test_graph_replay_equals_eager_steps():
    """Three optimiser steps replayed from a captured CUDA graph give the same parameters as three eager steps."""
    g = load_golden("c2_full_128_sub")
    case = golden_case(g)
    model, topo, params, batch, cfg = case
    B = params["pose"].shape[0]
    fid = torch.arange(B, device="cuda")
    oa, on = batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous()
    eager, _ = _fitter_from_case(case)
    for _ in range(5):
        eager.step(fid, oa, on, stage="both", update=True)
    graphed, _ = _fitter_from_case(case)
    replay = graphed.capture(fid, oa, on, stage="both", warmup=2)   # 2 warm-up steps + 1 captured-but-not-run
    for _ in range(3):
        replay()
    torch.cuda.synchronize()
    assert graphed.coarse.steps == 5
    # atomics reorder between runs, so compare to the tolerance of 5 Adam steps (lr 1e-3 / 1e-2)
    assert (graphed.coarse.data - eager.coarse.data).abs().max() < 2e-3
    assert (graphed.app.data - eager.app.data).abs().max() < 2e-2
    assert rel_l2(graphed.coarse.p["pose"].cpu(), eager.coarse.p["pose"].cpu()) < 1e-2


def _fitter_from_case(case, S=None):
    from harp_b200 import synth
    from harp_b200.step import HarpFitter
    model, topo, params, batch, cfg = case
    hc = _ctx(model, topo.edges0 is not None)
    B = params["pose"].shape[0]
    p = dict(params)
    p["light_positions"] = params["light_position"][None].repeat(B, 1)
    frames = dict(y_true=batch["y_true"], y_sil=batch["y_sil"], y_sil_eroded=batch["y_sil_eroded"], init_joints=batch["init_joints"])
    return HarpFitter(hc, p, frames, batch["ref_verts"], batch["uv_mask"], cfg["img_size"], cfg["focal_length"]), hc


@pytest.mark.parametrize("name", ["c2_full_128_sub", "c1_sil_256_coarse", "c4_arm_128_sub", "c2_full_512_sub", "arm_448_sub",
                                  "c4_arm_1024_sub"])
def test_fused_step_matches_reference_golden(name):
    """The whole fused step against vectors produced by the reference's own Python (tests/golden/make_golden.py)."""
    from harp_b200.step import LOSS_NAMES
    g = load_golden(name)
    case = golden_case(g)
    model, topo, params, batch, cfg = case
    ft, hc = _fitter_from_case(case)
    B = params["pose"].shape[0]
    fid = torch.arange(B, device="cuda")
    ft.step(fid, batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous(),
            stage=cfg["stage"], update=False)
    ld = ft.loss_dict()
    cnt = hc.counters()
    assert cnt["bin_overflow"] == 0 and cnt["sil_k_overflow_pixels"] == 0
    for k in LOSS_NAMES:
        if "loss_" + k in g:
            assert math.isclose(ld[k], float(g["loss_" + k]), rel_tol=2e-3, abs_tol=1e-7), (k, ld[k], float(g["loss_" + k]))
    assert math.isclose(ld["total"], float(g["loss_total"]), rel_tol=1e-3)
    d = (ft.last["alpha"].cpu() - torch.tensor(g["img_sil"])).abs()
    assert float((d <= 1e-4).double().mean()) >= 0.998
    grads = dict(ft.coarse.g); grads.update(ft.app.g)
    grads["rot"], grads["trans"], grads["wrist_pose"] = ft.last["d_rot"], ft.last["d_trans"], ft.last["d_wrist"]
    # Dense parameters: relative L2.  Texture / normal map: each hit pixel deposits +-w/(3 B S^2) on its own 4 texels
    # (sign of an L1 residual), so ONE pixel whose nearest face or residual sign differs moves the L2 norm by ~2/sqrt(N_hit)
    # (N_hit ~ 5e3 at 128^2); the faithful statement is element-wise agreement on all but a handful of texels.
    tol = dict(pose=1e-2, rot=1e-2, cam=1e-2, shape=1e-2, verts_disps=2e-2, texture=6e-2, normal_map=6e-2,
               light_positions=2e-2, amb_ratio=1e-2, trans=1e-2, wrist_pose=1e-2)
    for k, t in tol.items():
        if "grad_" + k not in g:
            continue
        gg = torch.tensor(g["grad_" + k])
        if float(gg.abs().max()) == 0.0:
            continue
        mine = grads[k].cpu().view(gg.shape)
        r = rel_l2(mine, gg)
        record(f"fused_step_vs_golden[{name}]", **{"grad_" + k: r})
        if k == "amb_ratio":
            # a scalar, heavily cancelling sum of +-w/(3 B S^2) * texel * (1 - cos*vis) * amb(1-amb) over the hit pixels'
            # channels.  A residual |rgb - y| at rounding level (a pixel whose prediction and clamped target are both white)
            # has sign 0 in one float32 evaluation and +-1 in another: the reference's own float32 run differs from the float64
            # oracle by 14 % on this scalar (profiles/r02_parity.json).  Bound: a few dozen such channel quanta.
            S_, B_ = int(g["S"]), int(g["B"])
            quantum = 2.0 / (3 * B_ * S_ * S_) / 4
            record(f"fused_step_vs_golden[{name}]", amb_ratio_abs=abs(float(mine) - float(gg)), amb_ratio_quanta=abs(float(mine) - float(gg)) / quantum)
            assert abs(float(mine) - float(gg)) <= 40 * quantum, (k, float(mine), float(gg))
            continue
        assert r < t, (k, r)
        if k in ("texture", "normal_map"):
            frac = frac_within(mine, gg, 1e-3 * float(gg.abs().max()))
            assert frac >= 0.9995, (k, frac)
