"""ORACLE — test infrastructure only.  Never imported by the product (`harp_b200/`); only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may use it, as the checker.

CPU restatement (pure torch, dtype-generic, autograd for gradients) of the HARP per-minibatch hot path
`optimize_sequence.py:446-579`: MANO LBS -> mesh build -> soft silhouette + shadowed PBR render -> losses.

PARITY STATUS: L2/L3 (this file) is PINNED against the reference's own, unmodified Python
(`manopth/manolayer.py`, `utils/visualize.py`, `renderer/renderer_helper.py`, `renderer/pbr_materials.py`,
`loss/*.py`) executed in the build container on the shim in `oracle/p3d_shim` — see `tests/golden/make_golden.py`
and `tests/test_oracle_golden.py`.  L1 (the PyTorch3D primitives inside `oracle/p3d_shim/pytorch3d`) is
PARITY-UNPINNED: pytorch3d==0.6.2 (`requirements.txt:56`) is un-vendored and not installable here; its published
algorithm is restated and self-checked by float64 finite differences.

Each function cites the reference lines it follows.
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "p3d_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from pytorch3d import _core as p3d  # noqa: E402
from pytorch3d import _raster  # noqa: E402

LOSS_WEIGHTS = {  # optimize_sequence.py:411-422 (vgg excluded: SURVEY.md section 2 row 12)
    "silhouette": 7.0, "kps_anchor": 10.0, "vert_disp_reg": 2.0, "normal": 0.1, "laplacian": 4.0, "arap": 0.2,
    "photo": 1.0, "albedo": 0.5, "normal_reg": 0.1,
}
SIL_SIGMA = 1e-7  # optimize_sequence.py:425-426
SIL_K = 50
SIL_BLUR = math.log(1.0 / 1e-4 - 1.0) * SIL_SIGMA  # renderer_helper.py:46
RGB_SIGMA, RGB_GAMMA = 1e-4, 1e-4  # BlendParams defaults, renderer_helper.py:67
SHADOW_RADIUS = 1.5  # renderer_helper.py:463
SHADOW_BIAS, SHADOW_SHARP = 0.008, 1000.0  # renderer_helper.py:402-406


# ------------------------------------------------------------------------------------------------- MANO LBS
def quat2mat(quat):
    """manopth/rodrigues_layer.py:15-40"""
    nq = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz, 2 * wz + 2 * xy,
                        w2 - x2 + y2 - z2, 2 * yz - 2 * wx, 2 * xz - 2 * wy, 2 * wx + 2 * yz,
                        w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def batch_rodrigues(axisang):
    """manopth/rodrigues_layer.py:43-54 (note the +1e-8 inside the norm)."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1).unsqueeze(-1)
    n = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * n], dim=1)
    return quat2mat(quat).view(-1, 9)


def mano_forward(m, pose_coeffs, betas, trans):
    """manopth/manolayer.py:108-296 for use_pca=False, axis-angle root, side='right', center_idx=None.
    m: dict of tensors named as the registered buffers (manolayer.py:70-97) + 'tips' (5 vertex ids, 270-273).
    Returns (verts_mm (B,778,3), joints_mm (B,21,3))."""
    B = pose_coeffs.shape[0]
    full_pose = torch.cat([pose_coeffs[:, :3], m["th_hands_mean"] + pose_coeffs[:, 3:48]], 1)  # 140-143
    rot_map = batch_rodrigues(full_pose.reshape(-1, 3)).view(B, 16 * 9)  # tensutils.py:6-12
    eye = torch.eye(3, dtype=rot_map.dtype).view(1, 9).repeat(B, 16)
    pose_map = (rot_map - eye)[:, 9:]
    root_rot = rot_map[:, :9].view(B, 3, 3)
    rot_map = rot_map[:, 9:]
    v_shaped = torch.matmul(m["th_shapedirs"], betas.transpose(1, 0)).permute(2, 0, 1) + m["th_v_template"]  # 185-187
    j = torch.matmul(m["th_J_regressor"], v_shaped)  # 188
    v_posed = v_shaped + torch.matmul(m["th_posedirs"], pose_map.transpose(0, 1)).permute(2, 0, 1)  # 191-192

    def with_zeros(t):  # tensutils.py:15-22
        pad = t.new_tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4).repeat(t.shape[0], 1, 1)
        return torch.cat([t, pad], 1)

    root_j = j[:, 0, :].contiguous().view(B, 3, 1)
    root_trans = with_zeros(torch.cat([root_rot, root_j], 2))  # 197-203
    all_rots = rot_map.view(B, 15, 3, 3)
    l1, l2, l3 = [1, 4, 7, 10, 13], [2, 5, 8, 11, 14], [3, 6, 9, 12, 15]
    r1, r2, r3 = all_rots[:, [i - 1 for i in l1]], all_rots[:, [i - 1 for i in l2]], all_rots[:, [i - 1 for i in l3]]
    j1, j2, j3 = j[:, l1], j[:, l2], j[:, l3]
    transforms = [root_trans.unsqueeze(1)]
    rel1 = with_zeros(torch.cat([r1, (j1 - root_j.transpose(1, 2)).unsqueeze(3)], 3).view(-1, 3, 4))
    root_flt = root_trans.unsqueeze(1).repeat(1, 5, 1, 1).view(B * 5, 4, 4)
    t1 = torch.matmul(root_flt, rel1); transforms.append(t1.view(B, 5, 4, 4))
    rel2 = with_zeros(torch.cat([r2, (j2 - j1).unsqueeze(3)], 3).view(-1, 3, 4))
    t2 = torch.matmul(t1, rel2); transforms.append(t2.view(B, 5, 4, 4))
    rel3 = with_zeros(torch.cat([r3, (j3 - j2).unsqueeze(3)], 3).view(-1, 3, 4))
    t3 = torch.matmul(t2, rel3); transforms.append(t3.view(B, 5, 4, 4))
    reorder = [0, 1, 6, 11, 2, 7, 12, 3, 8, 13, 4, 9, 14, 5, 10, 15]
    results = torch.cat(transforms, 1)[:, reorder]  # 241-242
    joint_js = torch.cat([j, j.new_zeros(B, 16, 1)], 2)
    tmp2 = torch.matmul(results, joint_js.unsqueeze(3))
    results2 = (results - torch.cat([tmp2.new_zeros(B, 16, 4, 3), tmp2], 3)).permute(0, 2, 3, 1)  # 245-247
    T = torch.matmul(results2, m["th_weights"].transpose(0, 1))  # 251
    rest_h = torch.cat([v_posed.transpose(2, 1), torch.ones((B, 1, v_posed.shape[1]), dtype=T.dtype)], 1)
    verts = (T * rest_h.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]  # 253-261
    jtr = results[:, :, :3, 3]
    jtr = torch.cat([jtr, verts[:, m["tips"]]], 1)  # 270-276
    jtr = jtr[:, [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]]  # 279
    if not bool(torch.norm(trans) == 0):  # 281-288 (center_idx None => the all-zero branch is a no-op, trans unused)
        jtr = jtr + trans.unsqueeze(1)
        verts = verts + trans.unsqueeze(1)
    return verts * 1000, jtr * 1000  # 291-292


# ------------------------------------------------------------------------------------------------- SMPL-X arm LBS
def smplxarm_forward(m, betas, global_orient, right_hand_pose, right_wrist_pose, transl):
    """hand_models_harp/body_models.py:2163-2390 (SMPLXARM.forward, return_type='mano_w_arm') for the arguments
    utils/visualize.py:37-40 passes; every other pose input is the layer's zero default (SMPL.__init__ 180-215, SMPLH
    638-660, SMPLXARM 2036-2098).  m: dict of tensors with the SMPLX npz field names + 'arm_vert', 'mano_joint',
    'extra_joints_idxs'.  Returns (verts_mm (B,1026,3), joints_mm (B,22,3))."""
    from smplx_shim import lbs as slbs
    B, dt = global_orient.shape[0], betas.dtype
    body_pose = torch.zeros(B, 63, dtype=dt)
    body_pose[:, 60:63] = right_wrist_pose  # 2299-2301
    z3, z45 = torch.zeros(B, 3, dtype=dt), torch.zeros(B, 45, dtype=dt)
    full_pose = torch.cat([global_orient.reshape(-1, 1, 3), body_pose.reshape(-1, 21, 3), z3.reshape(-1, 1, 3),
                           z3.reshape(-1, 1, 3), z3.reshape(-1, 1, 3), z45.reshape(-1, 15, 3),
                           right_hand_pose.reshape(-1, 15, 3)], dim=1).reshape(-1, 165)  # 2304-2311
    pose_mean = torch.cat([torch.zeros(75, dtype=dt), m["hands_meanl"], m["hands_meanr"]])  # create_mean_pose 2129-2147
    full_pose = full_pose + pose_mean  # 2315
    shape_components = torch.cat([betas, torch.zeros(B, 10, dtype=dt)], dim=-1)  # expression default zeros, 2323
    posedirs = m["posedirs"].reshape(-1, m["posedirs"].shape[-1]).T  # SMPL.__init__ 249-253
    parents = m["kintree_table"][0].long().clone(); parents[0] = -1
    vertices, joints = slbs.lbs(shape_components, full_pose, m["v_template"], m["shapedirs"], posedirs, m["J_regressor"],
                                parents, m["weights"])  # 2335-2339
    vertices = vertices - joints[:, None, 21, :]  # 2342
    joints = joints - joints[:, None, 21, :]
    joints = torch.cat([joints, torch.index_select(vertices, 1, m["extra_joints_idxs"])], dim=1)  # 2365 (landmarks are appended
    # after index 75 and never selected by joint_idx_tensor)
    joints = joints + transl.unsqueeze(1)  # 2378-2380
    vertices = vertices + transl.unsqueeze(1)
    return vertices[:, m["arm_vert"]] * 1000.0, joints[:, m["mano_joint"]] * 1000.0  # 2383-2385


# ------------------------------------------------------------------------------------------------- topology
class Topology:
    """Static tables: coarse faces -> PyTorch3D SubdivideMeshes layout (optimize_sequence.py:67-89), UVs."""

    def __init__(self, faces0, faces_uvs=None, verts_uvs=None, subdivide=True):
        faces0 = torch.as_tensor(faces0).long()
        V0 = int(faces0.max()) + 1
        m0 = p3d.Meshes(torch.zeros(1, V0, 3), faces0[None])
        self.V0, self.faces0 = V0, faces0
        if subdivide:
            sub = p3d.SubdivideMeshes(m0)
            self.edges0 = sub._edges
            self.faces = sub._subdivided_faces
            self.V = V0 + self.edges0.shape[0]
        else:
            self.edges0, self.faces, self.V = None, faces0, V0
        self.faces_uvs = None if faces_uvs is None else torch.as_tensor(faces_uvs).long()
        self.verts_uvs = None if verts_uvs is None else torch.as_tensor(verts_uvs)


def build_mesh(verts0_m, topo, verts_disps):
    """utils/visualize.py:45-64: subdivide, displace along the subdivided mesh's vertex normals."""
    B = verts0_m.shape[0]
    if topo.edges0 is not None:
        mids = verts0_m[:, topo.edges0].mean(dim=2)
        vs = torch.cat([verts0_m, mids], dim=1)
    else:
        vs = verts0_m
    if verts_disps is None:
        return vs
    faces = topo.faces[None].expand(B, -1, -1)
    mesh = p3d.Meshes(vs, faces)
    return vs + mesh.verts_normals_padded() * verts_disps.repeat(B, 1, 1)


# ------------------------------------------------------------------------------------------------- cameras
def cam_to_T(cam, focal, S):
    """utils/visualize.py:268 / renderer_helper.py:455"""
    return torch.stack([-cam[:, 1], -cam[:, 2], 2 * focal / (S * cam[:, 0] + 1e-9)], dim=1)


def cam_R(B, dtype):
    return torch.tensor([[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, 1.0]], dtype=dtype).repeat(B, 1, 1)


def project_ndc(verts, R, T, focal, S):
    """MeshRasterizer.transform with PerspectiveCameras(in_ndc=False), pp = S/2.
    Returns (B,V,3) = NDC x, NDC y, view-space z; and the view-space verts."""
    view = torch.bmm(verts, R) + T[:, None, :]
    z = view[..., 2:3]
    xy = (focal * view[..., :2] / z) / (S / 2.0)
    return torch.cat([xy, z], dim=-1), view


def light_camera(cam, light_positions, center, focal, S):
    """renderer_helper.py:454-468 process_info_for_shadow."""
    cT = cam_to_T(cam, focal, S)
    cR = cam_R(cam.shape[0], cam.dtype)
    lp = center + (light_positions - center) * (SHADOW_RADIUS / torch.linalg.norm(light_positions - center, dim=1, keepdim=True))
    lR = p3d.look_at_rotation(lp, at=center, up=((0, 1, 0),), device=lp.device)
    lT = -torch.bmm(lR.transpose(1, 2), lp[:, :, None])[:, :, 0]
    return lR, lT, cR, cT


# ------------------------------------------------------------------------------------------------- renders
def _frag(verts_ndc, faces, S, blur, K, clip):
    p2f, zb, ba, di, ov = _raster.rasterize_meshes(verts_ndc, faces, S, blur, K, True, clip)
    return p3d.Fragments(p2f, zb, ba, di), ov


def render_silhouette(verts, topo, cam, focal, S, K=SIL_K, sigma=SIL_SIGMA, return_overflow=False):
    """renderer_helper.py:37-58 silhouette renderer + visualize.py:258-285."""
    B = verts.shape[0]
    ndc, _ = project_ndc(verts, cam_R(B, verts.dtype), cam_to_T(cam, focal, S), focal, S)
    faces = topo.faces[None].expand(B, -1, -1)
    blur = math.log(1.0 / 1e-4 - 1.0) * sigma
    fr, ov = _frag(ndc, faces, S, blur, K, True)
    prob = torch.sigmoid(-fr.dists / sigma) * (fr.pix_to_face >= 0)
    alpha = 1.0 - torch.prod(1.0 - prob, dim=-1)
    return (alpha, ov) if return_overflow else alpha


def compute_tangent(n):
    """renderer/pbr_materials.py:58-77"""
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    s = (2 * (z >= 0)) - 1.0
    a = -1 / (s + z)
    b = x * y * a
    uv = torch.stack((1 + s * x * x * a, s * b, -s * x, b, s + y * y * a, -y), dim=-1)
    return uv.view(uv.shape[:-1] + (2, 3))


def sample_uv(maps, pixel_uvs):
    """TexturesUV.sample_textures for K=1. maps (B,H,W,C); pixel_uvs (B,S,S,2)."""
    tm = torch.flip(maps.permute(0, 3, 1, 2), [2])
    tex = F.grid_sample(tm, pixel_uvs * 2.0 - 1.0, mode="bilinear", align_corners=True, padding_mode="border")
    return tex.permute(0, 2, 3, 1)


def render_rgb_shadow(verts, topo, cam, light_positions, amb_ratio, texture, normal_map, focal, S, self_shadow=True,
                      return_aux=False):
    """optimize_sequence.py:474-488 -> MeshRendererShadow.forward (renderer_helper.py:331-412) ->
    SoftPhongShaderShadow (565-592) -> phong_shading_with_shadow (472-523) -> softmax_rgb_blend.
    texture / normal_map: (1,512,512,3) shared maps; normal_map is the *raw* parameter (normalised here as in
    visualize.py:98)."""
    B, dt = verts.shape[0], verts.dtype
    faces = topo.faces[None].expand(B, -1, -1)
    mesh = p3d.Meshes(verts, faces)
    vnorm = mesh.verts_normals_padded()
    lR, lT, cR, cT = light_camera(cam, light_positions, verts.mean(1), focal, S)
    ndc_c, _ = project_ndc(verts, cR, cT, focal, S)
    fr_c, _ = _frag(ndc_c, faces, S, 0.0, 1, False)
    mask = fr_c.pix_to_face >= 0  # (B,S,S,1)
    fp = mesh.faces_packed()
    interp = lambda attr: p3d.interpolate_face_attributes(fr_c.pix_to_face, fr_c.bary_coords, attr)[..., 0, :]
    pix_pos = interp(verts.reshape(-1, 3)[fp])  # (B,S,S,3)
    if self_shadow:
        ndc_l, _ = project_ndc(verts, lR, lT, focal, S)
        fr_l, _ = _frag(ndc_l, faces, S, 0.0, 1, False)
        flat = pix_pos.reshape(B, S * S, 3)
        pl = torch.bmm(flat, lR) + lT[:, None, :]  # light view space (renderer_helper.py:379)
        zl = pl[..., 2].reshape(B, S, S)
        # transform_points_screen (382): x_px = S/2 - (S/2) x_ndc
        xs = S / 2.0 - (S / 2.0) * ((focal * pl[..., 0] / pl[..., 2]) / (S / 2.0))
        ys = S / 2.0 - (S / 2.0) * ((focal * pl[..., 1] / pl[..., 2]) / (S / 2.0))
        xk = xs.round().long().reshape(B, S, S)
        yk = ys.round().long().reshape(B, S, S)
        bk = torch.arange(B).view(B, 1, 1).expand(B, S, S)
        zbl = fr_l.zbuf[..., 0]
        vis = torch.zeros_like(zl)
        for ii in (-1, 0, 1):
            for jj in (-1, 0, 1):
                d = zbl[bk, (yk + ii).clamp(0, S - 1), (xk + jj).clamp(0, S - 1)]
                vis = vis + torch.sigmoid((d - (zl - SHADOW_BIAS)) * SHADOW_SHARP)
        vis = vis / 9.0
    else:
        vis = torch.ones(B, S, S, dtype=dt)
    # shading
    pix_n = interp(vnorm.reshape(-1, 3)[fp])
    fuv = topo.verts_uvs.to(dt)[topo.faces_uvs]  # (F,3,2)
    pix_uv = interp(fuv.repeat(B, 1, 1))
    texel = sample_uv(texture.expand(B, -1, -1, -1), pix_uv)
    nm = sample_uv(F.normalize(normal_map, dim=-1).expand(B, -1, -1, -1), pix_uv)
    tb = compute_tangent(pix_n)
    new_n = F.normalize(-tb[..., 0, :] * nm[..., 0:1] - tb[..., 1, :] * nm[..., 1:2] + pix_n * nm[..., 2:3], dim=-1)
    amb = torch.sigmoid(amb_ratio)  # optimize_sequence.py:480
    direction = light_positions[:, None, None, :] - pix_pos
    nn_ = F.normalize(new_n, p=2, dim=-1, eps=1e-6)
    dd = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    ang = F.relu((nn_ * dd).sum(-1))
    diffuse = (1.0 - amb) * ang
    colors = (amb + diffuse * vis)[..., None] * texel  # specular_color = 0 (renderer_helper.py:440)
    # softmax_rgb_blend, K=1
    m = mask[..., 0].to(dt)
    prob = torch.sigmoid(-fr_c.dists[..., 0] / RGB_SIGMA) * m
    z_inv = (100.0 - fr_c.zbuf[..., 0]) / (100.0 - 1.0) * m
    z_inv_max = z_inv.clamp(min=1e-10)
    wnum = prob * torch.exp((z_inv - z_inv_max) / RGB_GAMMA)
    delta = torch.exp((1e-10 - z_inv_max) / RGB_GAMMA).clamp(min=1e-10)
    rgb = (wnum[..., None] * colors + delta[..., None] * 1.0) / (wnum + delta)[..., None]
    if return_aux:
        return rgb, dict(vis=vis, mask=mask[..., 0], pix_to_face=fr_c.pix_to_face[..., 0], zbuf=fr_c.zbuf[..., 0],
                         bary=fr_c.bary_coords[..., 0, :], light_zbuf=fr_l.zbuf[..., 0] if self_shadow else None,
                         light_p2f=fr_l.pix_to_face[..., 0] if self_shadow else None)
    return rgb


def _shade_common(verts, topo, fr, texture, normal_map, use_normal_map=True):
    """Per-fragment texels and (normal-mapped) normals for a (B,S,S,K) fragment set."""
    B = verts.shape[0]
    faces = topo.faces[None].expand(B, -1, -1)
    mesh = p3d.Meshes(verts, faces)
    fp = mesh.faces_packed()
    vnorm = mesh.verts_normals_padded()
    interp = lambda attr: p3d.interpolate_face_attributes(fr.pix_to_face, fr.bary_coords, attr)  # (B,S,S,K,D)
    pix_pos = interp(verts.reshape(-1, 3)[fp])
    pix_n = interp(vnorm.reshape(-1, 3)[fp])
    fuv = topo.verts_uvs.to(verts.dtype)[topo.faces_uvs].repeat(B, 1, 1)
    pix_uv = interp(fuv)
    K = fr.pix_to_face.shape[-1]
    S = fr.pix_to_face.shape[1]

    def sample(maps):
        uv = pix_uv.permute(0, 3, 1, 2, 4).reshape(B * K, S, S, 2)
        tm = torch.flip(maps.expand(B, -1, -1, -1).permute(0, 3, 1, 2), [2])[:, None].expand(-1, K, -1, -1, -1).reshape(B * K, 3, *maps.shape[1:3])
        t = F.grid_sample(tm, uv * 2.0 - 1.0, mode="bilinear", align_corners=True, padding_mode="border")
        return t.reshape(B, K, 3, S, S).permute(0, 3, 4, 1, 2)

    texel = sample(texture) if texture is not None else None
    if use_normal_map:
        nm = sample(F.normalize(normal_map, dim=-1))
        tb = compute_tangent(pix_n)
        pix_n = F.normalize(-tb[..., 0, :] * nm[..., 0:1] - tb[..., 1, :] * nm[..., 1:2] + pix_n * nm[..., 2:3], dim=-1)
    return pix_pos, pix_n, texel


def _softmax_rgb_blend(colors, fr, sigma=RGB_SIGMA, gamma=RGB_GAMMA, znear=1.0, zfar=100.0):
    """pytorch3d softmax_rgb_blend, background white."""
    mask = (fr.pix_to_face >= 0).to(colors.dtype)
    prob = torch.sigmoid(-fr.dists / sigma) * mask
    z_inv = (zfar - fr.zbuf) / (zfar - znear) * mask
    z_inv_max = torch.max(z_inv, dim=-1).values[..., None].clamp(min=1e-10)
    wnum = prob * torch.exp((z_inv - z_inv_max) / gamma)
    delta = torch.exp((1e-10 - z_inv_max) / gamma).clamp(min=1e-10)
    denom = wnum.sum(dim=-1)[..., None] + delta
    return ((wnum[..., None] * colors).sum(dim=-2) + delta * 1.0) / denom


def render_phong(verts, topo, cam, light_positions, texture, normal_map, focal, S):
    """Non-shadow colour renderer: get_renderers' phong_renderer (renderer_helper.py:60-83) = SoftPhongShaderPBR (106-144) +
    phong_shading_PBR (147-190): lights ambient .5 / diffuse .4 / specular .1, shininess 0 (visualize.py:259-263)."""
    B = verts.shape[0]
    ndc, _ = project_ndc(verts, cam_R(B, verts.dtype), cam_to_T(cam, focal, S), focal, S)
    fr, _ = _frag(ndc, topo.faces[None].expand(B, -1, -1), S, 0.0, 1, False)
    pix_pos, pix_n, texel = _shade_common(verts, topo, fr, texture, normal_map, True)
    direction = light_positions[:, None, None, None, :] - pix_pos
    ang = F.relu((F.normalize(pix_n, p=2, dim=-1, eps=1e-6) * F.normalize(direction, p=2, dim=-1, eps=1e-6)).sum(-1))
    colors = (0.5 + 0.4 * ang)[..., None] * texel + 0.1  # pow(alpha, shininess=0) == 1: constant specular
    return _softmax_rgb_blend(colors, fr)


def render_normals(verts, topo, cam, normal_map, focal, S, K=10):
    """Normal renderer: get_renderers' phong_normal_renderer (renderer_helper.py:85-101) = SoftPhongNormalShader (239-301) +
    phong_normal_shading (216-236), K=10 faces per pixel blended by softmax_rgb_blend."""
    B = verts.shape[0]
    ndc, _ = project_ndc(verts, cam_R(B, verts.dtype), cam_to_T(cam, focal, S), focal, S)
    fr, _ = _frag(ndc, topo.faces[None].expand(B, -1, -1), S, 0.0, K, False)
    _, pix_n, _ = _shade_common(verts, topo, fr, None, normal_map, normal_map is not None)
    colors = (pix_n * torch.tensor([1.0, -1.0, -1.0], dtype=verts.dtype) + 1.0) / 2.0
    return _softmax_rgb_blend(colors, fr)


# ------------------------------------------------------------------------------------------------- losses
def kps_loss(gt_kps, pred_kps_m):
    """loss/kps_loss.py:4-17 (use_arm=False)."""
    gt = gt_kps - gt_kps[:, 0, None, :]
    pr = (pred_kps_m - pred_kps_m[:, 0, None, :]) * 1000.0
    return torch.mean((torch.norm(gt - pr, dim=2) / 100.0) ** 2.0)


def arap_loss(verts, ref_verts, topo):
    """loss/arap.py:4-57; ref_verts (1,V,3) constant."""
    B = verts.shape[0]
    m = p3d.Meshes(verts[:1].detach(), topo.faces[None])
    e = m.edges_packed()
    v0, v1 = verts[:, e[:, 0]], verts[:, e[:, 1]]
    r0, r1 = ref_verts[:, e[:, 0]], ref_verts[:, e[:, 1]]
    loss = ((v0 - v1).norm(dim=2, p=2) * 1000.0 - (r0 - r1).norm(dim=2, p=2) * 1000.0) ** 2.0
    return (loss * (1.0 / e.shape[0])).sum() / B


def mesh_reg_losses(verts, topo):
    B = verts.shape[0]
    mesh = p3d.Meshes(verts, topo.faces[None].expand(B, -1, -1))
    return p3d.mesh_laplacian_smoothing(mesh), p3d.mesh_normal_consistency(mesh)


def smooth_texture_reg(tex, offsets, uv_mask):
    """loss/texture_reg.py:5-30 / 48-66 with the CPU-RNG neighbour offsets passed in."""
    t = tex.squeeze(0)
    H, W = t.shape[:2]
    gx, gy = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    tx = torch.clamp(gx + offsets[:, :, 0].long(), 0, H - 1)
    ty = torch.clamp(gy + offsets[:, :, 1].long(), 0, W - 1)
    diff = torch.norm(t - t[tx, ty], p=1, dim=2) / 3.0
    if uv_mask is not None:
        diff = diff * uv_mask
    return diff.mean()


def normal_reg(normal_map, offsets, uv_mask):
    """loss/texture_reg.py:33-45: note norm over dim=2 of the 4-D map (= W axis), bug-compatible."""
    z = torch.tensor([0.0, 0.0, 1.0], dtype=normal_map.dtype)
    close = (torch.norm(normal_map - z, p=2, dim=2) / 3.0).mean()
    return 0.2 * close + smooth_texture_reg(normal_map, offsets, uv_mask)


# ------------------------------------------------------------------------------------------------- the step
def forward_losses(model, topo, params, batch, cfg):
    """One minibatch of optimize_sequence.py:446-561.
    params: pose (B,45) rot (B,3) trans (B,3) cam (B,3) [per-frame rows already gathered by fid], shape (10,),
            verts_disps (V,1), texture, normal_map (1,512,512,3), light_position (3,), amb_ratio ().
    batch: y_true (B,S,S,3), y_sil (B,S,S), y_sil_eroded (B,S,S), init_joints (B,21,3) mm, ref_verts (1,V,3) m,
           uv_mask (512,512), off_albedo / off_normal (512,512,2) int.
    cfg: img_size, focal_length, stage in {'coarse','both','app'} (507-515), self_shadow.
    Returns (dict of unweighted losses, dict of images)."""
    S, focal = cfg["img_size"], cfg["focal_length"]
    B = params["pose"].shape[0]
    stage = cfg.get("stage", "both")
    if cfg.get("use_arm", False):  # visualize.py:36-40
        verts_mm, joints_mm = smplxarm_forward(model, params["shape"].repeat([B, 1]), params["rot"], params["pose"],
                                               params["wrist_pose"], params["trans"])
    else:
        verts_mm, joints_mm = mano_forward(model, torch.cat((params["rot"], params["pose"]), 1),
                                           params["shape"].repeat([B, 1]), params["trans"])  # visualize.py:42-44
    verts0, joints = verts_mm / 1000.0, joints_mm / 1000.0
    verts = build_mesh(verts0, topo, params.get("verts_disps"))
    light = params["light_position"][None].repeat(B, 1)  # share_light_position (453-454)
    out, img = {}, {}
    sil, ov = render_silhouette(verts, topo, params["cam"], focal, S, return_overflow=True)
    img["sil"], img["sil_overflow"] = sil, ov
    if cfg.get("render_rgb", True):
        img["rgb"] = render_rgb_shadow(verts, topo, params["cam"], light, params["amb_ratio"], params["texture"],
                                       params["normal_map"], focal, S, cfg.get("self_shadow", True))
    if stage in ("coarse", "both"):
        out["silhouette"] = F.l1_loss(batch["y_sil"], sil)
        if "init_joints" in batch:
            out["kps_anchor"] = kps_loss(batch["init_joints"], joints[:, :21] if cfg.get("use_arm", False) else joints)
        if params.get("verts_disps") is not None and cfg.get("mesh_losses", True):
            out["vert_disp_reg"] = torch.sum(params["verts_disps"] ** 2.0)
            out["laplacian"], out["normal"] = mesh_reg_losses(verts, topo)
            out["arap"] = arap_loss(verts, batch["ref_verts"], topo)
    if stage in ("both", "app") and cfg.get("render_rgb", True):
        me = batch["y_sil_eroded"].unsqueeze(-1)
        out["photo"] = F.l1_loss(batch["y_true"] * me, img["rgb"] * me)
        out["albedo"] = smooth_texture_reg(params["texture"], batch["off_albedo"], batch["uv_mask"])
        out["normal_reg"] = normal_reg(params["normal_map"], batch["off_normal"], batch["uv_mask"])
    img["verts"], img["joints"] = verts, joints
    return out, img


def total_loss(losses):
    s = 0.0
    for k, v in losses.items():
        s = s + v * LOSS_WEIGHTS[k]
    return s


GRAD_KEYS = ("pose", "rot", "trans", "wrist_pose", "cam", "shape", "verts_disps", "texture", "normal_map", "light_position",
             "amb_ratio")


def step(model, topo, params, batch, cfg):
    """forward + backward (optimize_sequence.py:556-569). Returns (losses, images, grads)."""
    p = {k: (v.detach().clone().requires_grad_(True) if k in GRAD_KEYS and v is not None else v) for k, v in params.items()}
    losses, img = forward_losses(model, topo, p, batch, cfg)
    tot = total_loss(losses)
    keys = [k for k in GRAD_KEYS if k in p and p[k] is not None]
    gs = torch.autograd.grad(tot, [p[k] for k in keys], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(keys, gs)}
    losses = {k: v.detach() for k, v in losses.items()}
    losses["total"] = tot.detach()
    return losses, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in img.items()}, grads


# ------------------------------------------------------------------------------------------------- upstream MANO fit
def mano_fit(model, pred_vertices, epoch_coarse=500, epoch_fine=700, max_tries=4):
    """metro_modifications/hand_utils.py:16-133 (optimize_for_mano_param) on CPU: 500 Adam steps on [rot, trans] at
    lr 1e-1 (75-91), 700 on [rot, pose, shape, trans] at lr 1e-2 (98-106) of MSELoss(ManoLayer verts mm, 1000 * pred) (74),
    restarted while the last loss exceeds 10 (54, 108-114).  PINNED: `tests/golden/fit_mano_b2.npz` holds the outputs of the
    reference's own function run in the build container (`tests/golden/make_golden_fit.py`).
    Returns (dict of tensors: joints, verts [mm], rot, pose, shape, trans; last loss)."""
    B = pred_vertices.shape[0]
    target = pred_vertices.detach() * 1000.0  # 39
    target_mean = target.mean(1)  # 40
    mse = torch.nn.MSELoss()
    loss = None
    for try_num in range(max_tries):  # 54
        shape = torch.zeros(B, 10, requires_grad=True)  # 57-58
        rot = torch.zeros(B, 3, requires_grad=True)  # 59-60
        pose = torch.zeros(B, 45, requires_grad=True)  # 61-62
        trans = (torch.zeros(B, 3) + target_mean / 1000.0).requires_grad_()  # 63-64
        for group, lr, n in (([rot, trans], 1e-1, epoch_coarse), ([rot, pose, shape, trans], 1e-2, epoch_fine)):
            opt = torch.optim.Adam(group, lr=lr)  # 75 / 98
            for _ in range(n):
                verts, _ = mano_forward(model, torch.cat((rot, pose), 1), shape, trans)
                loss = mse(verts, target)
                opt.zero_grad()
                loss.backward()
                opt.step()
        if float(loss) <= 10.0:  # 113
            break
    with torch.no_grad():
        verts, joints = mano_forward(model, torch.cat((rot, pose), 1), shape, trans)  # 120
    return dict(joints=joints, verts=verts, rot=rot.detach(), pose=pose.detach(), shape=shape.detach(), trans=trans.detach()), float(loss)


# ------------------------------------------------------------------------------------------------- evaluation metrics, resume
def fill_bg(img, mask):
    """utils/eval_util.py:28-31."""
    img = img * mask.unsqueeze(-1)
    return img + (mask.unsqueeze(-1) - 1) * -1.0


def l1_diff(ref_image, ref_mask, pred_image, pred_mask):
    """utils/eval_util.py:34-39 (the masks are unused by the reference)."""
    return torch.mean(torch.abs(ref_image - pred_image))


def sil_iou(ref_masks, pred_masks):
    """utils/eval_util.py:41-49.  PINNED with l1_diff / fill_bg: tests/golden/eval_metrics.npz (make_golden_misc.py)."""
    r, p = ref_masks >= 0.5, pred_masks >= 0.5
    return torch.mean(torch.logical_and(r, p).sum([1, 2]) / torch.logical_or(r, p).sum([1, 2]))


def resume_params(params):
    """optimize_sequence.py:366-389, statement by statement (inline code of main(); restated, parity unpinned)."""
    params = dict(params)
    temp_pose = params["pose"].detach().clone()
    for i in range(params["pose"].shape[0] // 30 - 1):
        for j in range(30):
            temp_pose[i * 30 + j] = ((30 - j) * params["pose"][i * 30] + j * params["pose"][i * 30 + 30]) / 30.0
    params["pose"] = temp_pose
    params["trans"] = torch.zeros_like(params["trans"]) + params["trans"].mean(0)
    params["rot"] = torch.zeros_like(params["rot"]) + params["rot"].mean(0)
    if "wrist_pose" not in params:
        params["wrist_pose"] = torch.zeros([params["pose"].shape[0], 3])
    if "amb_ratio" not in params:
        params["amb_ratio"] = torch.tensor(0.4)
    if "normal_map" not in params:
        params["normal_map"] = torch.tensor([0.0, 0.0, 1.0]).repeat(1, 512, 512, 1)
    return params


# ------------------------------------------------------------------------------------------------- upstream smoothing
def _smooth_l2(x):  # hand_utils.py:460
    return torch.sum(x ** 2)


def _smooth_poses(x, n_frames):  # LossSmoothPoses(1, nFrames).poses, hand_utils.py:505-513
    interp = x.clone().detach()
    interp[1:-1] = (interp[1:-1] + interp[:-2] + interp[2:]) / 3
    return _smooth_l2(x[1:-1] - interp[1:-1]) / (n_frames - 2)


def _kp3d(est, target, n_frames):  # LossKeypoints3D / LossAnchor .loss_func, hand_utils.py:472-496
    nj = min([est.shape[1], target.shape[1], 21])
    return _smooth_l2(est[:, :nj, :3] - target[:, :nj, :3]) / n_frames


def smooth_seq(model, params_in, img_res=224, total_iter_pose=1000, total_iter_cam=1000, forward=None):
    """metro_modifications/hand_utils.py:540-688 (optimize_smooth_seq, MANO branch) on CPU.
    Stage 1 (562-603): Adam(lr 1e-3) over [rot, pose, shape, cam] of 1e-2 * k3d(root-relative joints vs the input's) +
    1e-1 * temporal smoothness of the root-relative joints, stopped when the running loss stops falling by 1e-5 (592-595).
    Stage 2 (620-683): Adam(lr 1e-3, ReduceLROnPlateau patience 10) over cam of anchor + smoothness of the camera-relative
    root (1e-2 each).  Note the reference's broadcast at 643 / 665: camera_t is (T,1,3), the root joints (T,3), so
    `cam_rel_root` is (T,T,3) — reproduced.  PINNED: tests/golden/smooth_seq_t12.npz (make_golden_smooth.py).
    `forward(pose48, shape, trans) -> (verts_mm, joints_mm)` defaults to the oracle's mano_forward."""
    fwd = forward or (lambda p, s, t: mano_forward(model, p, s, t))
    p = {k: torch.as_tensor(v).float().clone() for k, v in params_in.items()}
    p["cam"] = p["cam"].unsqueeze(1)  # 547
    T = p["pose"].shape[0]
    target = p["joints"] - p["joints"][:, 0, :].unsqueeze(1)  # LossKeypoints3D.__init__, 463-464
    opt_params = [p["rot"], p["pose"], p["shape"], p["cam"]]
    for q in opt_params:
        q.requires_grad = True
    opt = torch.optim.Adam(opt_params, lr=1e-3)
    prev_loss = 999999.0
    for it in range(total_iter_pose):
        _, joints = fwd(torch.cat((p["rot"], p["pose"]), 1), p["shape"], p["trans"])
        rel = joints - joints[:, 0, :].unsqueeze(1)
        loss = _kp3d(rel, target, T) * 1e-2 + _smooth_poses(rel, T) * 1e-1  # smooth_body has weight 0 (568)
        if it > 0 and prev_loss - loss.item() < 0.00001:
            break
        prev_loss = (prev_loss + loss.item()) / 2.0
        opt.zero_grad()
        loss.backward()
        opt.step()
    with torch.no_grad():
        verts, joints = fwd(torch.cat((p["rot"], p["pose"]), 1), p["shape"], p["trans"])
    p["joints"], p["verts"] = joints, verts
    focal = 1000.0 * (img_res / 224.0)
    root = joints[:, 0, :] / 1000.0

    def cam_rel(cam):
        camera_t = torch.stack([cam[:, :, 1], cam[:, :, 2], 2 * focal / (img_res * cam[:, :, 0] + 1e-9)], dim=2)
        return camera_t + root  # (T,1,3) + (T,3) -> (T,T,3)

    anchor = cam_rel(p["cam"]).detach()
    opt = torch.optim.Adam([p["cam"]], lr=1e-3)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=10)
    for it in range(total_iter_cam):
        crr = cam_rel(p["cam"])
        loss = _kp3d(crr, anchor, T) * 1e-2 + _smooth_poses(crr, T) * 1e-2
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step(loss)
    p["cam"] = p["cam"].squeeze(1)
    return {k: v.detach() for k, v in p.items()}


# ------------------------------------------------------------------------------------------------- turntable
def turntable_verts(verts):
    """utils/visualize.py:162-196: the 72 vertex sets of render_360 — 36 cumulative 10-degree turns about Y, then 36 about X,
    around the centroid of the ORIGINAL mesh, with pytorch3d.transforms.RotateAxisAngle (restated in the shim)."""
    center = torch.mean(verts, dim=1, keepdim=True)  # 160
    out = []
    for axis in ("Y", "X"):
        for _ in range(36):
            verts = p3d.RotateAxisAngle(10, axis).transform_points(verts - center) + center
            out.append(verts)
    return out
