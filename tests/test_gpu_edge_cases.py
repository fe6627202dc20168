"""GPU parity of the fused step (tile_raster + shade_deferred, the optimised path) on framings the golden cases do not reach:
a hand partly outside the image, a frame with no hand at all, a hand that crosses the camera plane (vertices with z <= 0:
the classification's exact path), a tiny hand (many faces per tile: multi-chunk tiles and the top-K deferral), and an odd
batch size.  The checker is the CPU oracle's whole step (`O.step`) on identical inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import harp_oracle as O
from util import load_golden, golden_case, rel_l2, record
from test_gpu_kernels import _fitter_from_case


def _variant(kind, dtype=torch.float32):
    g = load_golden("c2_full_128_sub")
    model, topo, params, batch, cfg = golden_case(g, dtype=dtype)
    params = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in params.items()}
    batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    cam = params["cam"]
    if kind == "partly_outside":
        cam[:, 1] += 0.09          # about half of the hand leaves the image on one side
        cam[1, 2] -= 0.07
    elif kind == "one_empty_frame":
        cam[1, 1] += 5.0           # frame 1: nothing in view at all
    elif kind == "crosses_camera_plane":
        cam[0, 0] = 160.0          # t_z = 2 f / (S s) ~ 5.6 cm: part of the hand lies behind the camera (z <= 0)
    elif kind == "tiny_hand":
        cam[:, 0] = 1.5            # ~15 px wide at 128^2: tens of faces per pixel, hundreds per tile
    elif kind == "batch_of_three":
        idx = torch.tensor([0, 1, 0])
        for k in ("pose", "rot", "trans", "cam"):
            params[k] = params[k][idx].clone()
        params["cam"][2, 1] += 0.03
        for k in ("y_true", "y_sil", "y_sil_eroded", "init_joints"):
            batch[k] = batch[k][idx].clone()
    return model, topo, params, batch, cfg


@pytest.mark.parametrize("kind", ["partly_outside", "one_empty_frame", "crosses_camera_plane", "tiny_hand", "batch_of_three"])
def test_fused_step_edge_framing(kind):
    case = _variant(kind)
    model, topo, params, batch, cfg = case
    ref_losses, ref_img, ref_grads = O.step(model, topo, params, batch, cfg)
    ft, hc = _fitter_from_case(case)
    B = params["pose"].shape[0]
    fid = torch.arange(B, device="cuda")
    ft.step(fid, batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous(),
            stage=cfg["stage"], update=False)
    ld = ft.loss_dict()
    cnt = hc.counters()
    assert cnt["bin_overflow"] == 0 and cnt["sil_k_overflow_pixels"] == 0, cnt
    assert all(math.isfinite(v) for v in ld.values()), ld
    assert torch.isfinite(ft.coarse.grad).all() and torch.isfinite(ft.app.grad).all()
    # images: alpha is exactly 0/1 outside the +-1/3 px edge bands (tolerances as tests/test_gpu_kernels.py)
    da = (ft.last["alpha"].cpu() - ref_img["sil"]).abs()
    assert float((da <= 1e-4).double().mean()) >= 0.998, float((da <= 1e-4).double().mean())
    drgb = (ft.last["rgb"].cpu() - ref_img["rgb"]).abs().amax(-1)
    assert float((drgb <= 1e-4).double().mean()) >= 0.995, float((drgb <= 1e-4).double().mean())
    if kind == "one_empty_frame":
        assert float(ft.last["alpha"][1].abs().max()) == 0.0 and float((ft.last["rgb"][1] - 1.0).abs().max()) == 0.0
    for k in ("silhouette", "photo", "laplacian", "arap", "kps_anchor"):
        assert math.isclose(ld[k], float(ref_losses[k]), rel_tol=3e-3, abs_tol=1e-7), (k, ld[k], float(ref_losses[k]))
    assert math.isclose(ld["total"], float(ref_losses["total"]), rel_tol=2e-3)
    grads = dict(ft.coarse.g)
    errs = {k: rel_l2(grads[k].cpu().view(ref_grads[k].shape), ref_grads[k]) for k in ("pose", "cam", "shape", "verts_disps")}
    rec = dict(alpha_frac_1e4=float((da <= 1e-4).double().mean()), rgb_frac_1e4=float((drgb <= 1e-4).double().mean()), grads=errs)
    if kind == "crosses_camera_plane":
        # vertices with z ~ 0 project to coordinates of order 1/z: the float32 evaluation of THIS problem is itself only
        # ~4e-2 accurate (the float32 oracle against the float64 oracle, measured below), so the statement is: the CUDA path
        # is as close to the float64 value as the reference's float32 arithmetic is
        _, _, g64 = O.step(*_variant(kind, torch.float64))
        e_o32 = {k: rel_l2(ref_grads[k], g64[k]) for k in errs}
        e_cu = {k: rel_l2(grads[k].cpu().view(g64[k].shape), g64[k]) for k in errs}
        rec.update(o32_vs_o64=e_o32, cuda_vs_o64=e_cu)
        record(f"edge_framing[{kind}]", **rec)
        for k in errs:
            assert e_cu[k] <= 1.5 * e_o32[k] + 1e-3, (k, e_cu[k], e_o32[k])
        return
    record(f"edge_framing[{kind}]", **rec)
    # every framing meets north_star's rtol 1e-3 against the float32 oracle on identical inputs
    for k, e in errs.items():
        assert e < 1e-3, (k, e)


@pytest.mark.parametrize("K", [3, 8])
def test_fused_step_top_k_deferral_matches_standalone_silhouette(K):
    """faces_per_pixel smaller than the candidates on a pixel: the fused pixel pass must defer those pixels to the exact
    top-K kernel exactly as the stand-alone silhouette op does (which tests/test_gpu_kernels.py pins to the oracle at K=3, 8)."""
    from harp_b200.ops import Silhouette
    g = load_golden("c2_full_128_sub")
    case = golden_case(g)
    model, topo, params, batch, cfg = case
    ft, hc = _fitter_from_case(case)
    ft.faces_per_pixel = K
    B = params["pose"].shape[0]
    fid = torch.arange(B, device="cuda")
    ft.step(fid, batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous(),
            stage="both", update=False)
    alpha_fused = ft.last["alpha"].clone()
    d_cam_fused = ft.coarse.g["cam"].clone()
    loss_fused = ft.loss_dict()["silhouette"]
    S, focal = cfg["img_size"], cfg["focal_length"]
    verts = ft.last["verts"].detach().clone()
    cam = params["cam"].cuda().requires_grad_(True)
    a = Silhouette.apply(hc, verts, O.cam_R(B, torch.float32).cuda(), O.cam_to_T(cam, focal, S), focal, S, 1e-7, K)
    assert float((a - alpha_fused).abs().max()) <= 1e-6
    ref50 = Silhouette.apply(hc, verts, O.cam_R(B, torch.float32).cuda(), O.cam_to_T(cam, focal, S), focal, S, 1e-7, 50)
    assert float((ref50 - a).abs().max()) > 1e-3, "K must change the image for this test to mean anything"
    loss = (a - batch["y_sil"].cuda()).abs().mean()
    assert math.isclose(float(loss), loss_fused, rel_tol=1e-5)


@pytest.mark.parametrize("stage", ["coarse", "app"])
def test_fused_step_stage_gating(stage):
    """COARSE_OPT-only and APP_OPT-only epochs (optimize_sequence.py:507-515): the other stage's losses and gradients must be
    absent, the active ones must match the oracle."""
    g = load_golden("c2_full_128_sub")
    model, topo, params, batch, cfg = golden_case(g)
    cfg = dict(cfg, stage=stage)
    ref_losses, _, ref_grads = O.step(model, topo, params, batch, cfg)
    ft, hc = _fitter_from_case((model, topo, params, batch, cfg))
    B = params["pose"].shape[0]
    ft.step(torch.arange(B, device="cuda"), batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous(),
            stage=stage, update=False)
    ld = ft.loss_dict()
    for k in ("silhouette", "photo", "laplacian", "arap", "kps_anchor", "albedo", "normal_reg"):
        if k in ref_losses:
            assert math.isclose(ld[k], float(ref_losses[k]), rel_tol=2e-3, abs_tol=1e-7), (k, ld[k], float(ref_losses[k]))
        else:
            assert ld[k] == 0.0, (k, ld[k])
    grads = dict(ft.coarse.g); grads.update(ft.app.g)
    for k, t in dict(pose=1e-2, cam=1e-2, shape=1e-2, verts_disps=2e-2, texture=6e-2, normal_map=6e-2).items():
        gg = ref_grads[k]
        mine = grads[k].cpu().view(gg.shape)
        if float(gg.abs().max()) == 0.0:
            assert float(mine.abs().max()) == 0.0, k
        else:
            assert rel_l2(mine, gg) < t, (k, rel_l2(mine, gg))


def test_abi_error_behaviour_on_the_device():
    """Errors are negative return codes with a message (harp_last_error), never a crash or a silent fallback: missing
    reservation, an unsupported mode, a null tensor, a CPU tensor, a context without a skinning model / UV tables."""
    import ctypes as C
    from harp_b200 import _lib, synth
    from harp_b200.ops import HarpContext
    L = _lib.lib()
    tmpl = synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    hc = HarpContext(pack["th_faces"], subdivide=True, faces_uvs=tmpl["faces_uvs"], verts_uvs=tmpl["verts_uvs"], mano=pack)
    B, S = 2, 64
    verts = torch.zeros(B, hc.V, 3, device="cuda")
    R = torch.eye(3, device="cuda").repeat(B, 1, 1).contiguous()
    T = torch.tensor([[0.0, 0.0, 1.0]], device="cuda").repeat(B, 1).contiguous()
    alpha = torch.empty(B, S, S, device="cuda")
    st = _lib.stream()
    f = C.c_float
    args = lambda mode, a: (hc.h, mode, B, S, f(300.0), f(1e-7), 50, _lib.ptr(verts), _lib.ptr(R), _lib.ptr(T), a, None, f(0.0), None,
                            None, None, None, st)
    rc = L.harp_silhouette(*args(0, _lib.ptr(alpha)))          # no harp_ctx_reserve(B, S) yet
    assert rc < 0 and b"reserve" in L.harp_last_error()
    hc.reserve(B, S)
    assert L.harp_silhouette(*args(0, _lib.ptr(alpha))) == 0
    rc = L.harp_silhouette(*args(7, _lib.ptr(alpha)))
    assert rc < 0 and b"mode" in L.harp_last_error()
    rc = L.harp_silhouette(*args(0, None))
    assert rc < 0 and b"alpha" in L.harp_last_error()
    with pytest.raises(_lib.HarpError):
        _lib.ptr(torch.zeros(3))                               # CPU tensor: there is no CPU path
    with pytest.raises(_lib.HarpError):
        _lib.check(rc)
    bare = HarpContext(pack["th_faces"], subdivide=False)      # topology only: no skinning model, no UVs
    out_v, out_j = torch.empty(B, 778, 3, device="cuda"), torch.empty(B, 21, 3, device="cuda")
    z = torch.zeros(B, 48, device="cuda")
    rc = L.harp_lbs_mano_fwd(bare.h, B, _lib.ptr(z), _lib.ptr(z[:, :10].contiguous()), 0, _lib.ptr(z[:, :3].contiguous()),
                             _lib.ptr(out_v), _lib.ptr(out_j), st)
    assert rc < 0 and b"MANO" in L.harp_last_error()
    torch.cuda.synchronize()


def test_learning_rate_change_under_a_captured_graph():
    """The reference steps two ReduceLROnPlateau schedulers (optimize_sequence.py:312-321, 601-603).  The learning rates live
    on the device, so assigning `lr_coarse` / `lr_app` changes the update of an already captured step."""
    g = load_golden("c2_full_128_sub")
    case = golden_case(g)
    model, topo, params, batch, cfg = case
    ft, hc = _fitter_from_case(case)
    B = params["pose"].shape[0]
    fid = torch.arange(B, device="cuda")
    oa, on = batch["off_albedo"].cuda().int().contiguous(), batch["off_normal"].cuda().int().contiguous()
    replay = ft.capture(fid, oa, on, stage="both", warmup=1)
    replay()
    torch.cuda.synchronize()
    before_c, before_a = ft.coarse.data.clone(), ft.app.data.clone()
    ft.lr_coarse, ft.lr_app = 0.0, 0.0            # a scheduler driving the rates to zero: parameters must freeze
    replay()
    torch.cuda.synchronize()
    assert torch.equal(ft.coarse.data, before_c) and torch.equal(ft.app.data, before_a)
    ft.lr_coarse, ft.lr_app = 1e-3, 1e-2
    replay()
    torch.cuda.synchronize()
    assert float((ft.coarse.data - before_c).abs().max()) > 1e-5 and float((ft.app.data - before_a).abs().max()) > 1e-4
    assert ft.lr_coarse == 1e-3 and ft.lr_app == 1e-2


def test_vertex_normals_on_a_high_valence_mesh_use_the_csr_tables():
    """The hand / arm meshes have at most 8 faces around a vertex, so their gathers read the fixed-stride ring table; a topology with
    a higher valence falls back to the CSR tables.  Both paths must give PyTorch3D's area-weighted vertex normals and their
    gradient (Meshes._compute_vertex_normals)."""
    import numpy as np
    from harp_b200.ops import HarpContext, VertexNormals
    for n_fan in (7, 11):  # a closed double fan around an axis: the two poles have valence n_fan
        ring = [(np.cos(2 * np.pi * k / n_fan), np.sin(2 * np.pi * k / n_fan), 0.1 * np.sin(3.0 * k)) for k in range(n_fan)]
        verts = np.array([(0.0, 0.0, 0.7), (0.0, 0.0, -0.6)] + ring, np.float32)
        faces = []
        for k in range(n_fan):
            a, b = 2 + k, 2 + (k + 1) % n_fan
            faces += [(0, a, b), (1, b, a)]
        faces = np.array(faces, np.int32)
        hc = HarpContext(faces, subdivide=False, device="cuda:0")
        v = torch.tensor(verts[None], device="cuda", requires_grad=True)
        n = VertexNormals.apply(hc, v)
        w = torch.linspace(0.5, 1.5, n.numel(), device="cuda").view_as(n)
        (n * w).sum().backward()
        vr = torch.tensor(verts[None], dtype=torch.float64, requires_grad=True)
        f = torch.tensor(faces, dtype=torch.long)
        p0, p1, p2 = vr[0, f[:, 0]], vr[0, f[:, 1]], vr[0, f[:, 2]]
        acc = torch.zeros(len(verts), 3, dtype=torch.float64)
        acc = acc.index_add(0, f[:, 0], torch.cross(p1 - p0, p2 - p0, dim=1))
        acc = acc.index_add(0, f[:, 1], torch.cross(p2 - p1, p0 - p1, dim=1))
        acc = acc.index_add(0, f[:, 2], torch.cross(p0 - p2, p1 - p2, dim=1))
        nr = torch.nn.functional.normalize(acc, eps=1e-6, dim=1)[None]
        (nr * w.double().cpu()).sum().backward()
        assert float((n.detach().cpu().double() - nr.detach()).abs().max()) < 2e-6, n_fan
        assert float((v.grad.cpu().double() - vr.grad).norm() / vr.grad.norm()) < 1e-5, n_fan
