"""Synthetic fitting problems on the real hand-template topology, built with the CUDA path.

`make_scene` returns everything `HarpFitter` needs for a T-frame sequence at S x S: initial parameters laid out like
`optimize_sequence.py:init_params` (181-250), target images rendered by this package's own forward kernels from a
perturbed parameter set (pose noise, camera shift, different texture seed), masks (alpha >= 0.5, 3x3 erosion x2 as
`utils/data_util.py:18-20`), initial-joint anchors and the frame-0 ARAP reference mesh (`optimize_sequence.py:429-435`).
"""
import torch

from . import camera, synth
from .body_models import SMPLXARM
from .ops import HarpContext, Lbs, LbsMano, MeshBuild, Shade, Silhouette


def make_context(device="cuda:0", subdivide=True, seed=0, model="hand"):
    """model 'hand': MANO-shaped synthetic pack (778 -> 3093 verts); 'arm': SMPL-X-arm-shaped pack (1026 -> 4083 verts)."""
    tmpl = synth.load_template(model)
    if model == "arm":
        pack = synth.make_smplx_pack(seed=seed, template=tmpl)
        hc = HarpContext(tmpl["faces0"], subdivide=subdivide, faces_uvs=tmpl["faces_uvs"], verts_uvs=tmpl["verts_uvs"],
                         lbs=synth.reduce_smplx_arm(pack, tmpl), device=device)
    else:
        pack = synth.make_mano_pack(seed=seed, template=tmpl)
        hc = HarpContext(pack["th_faces"], subdivide=subdivide, faces_uvs=tmpl["faces_uvs"], verts_uvs=tmpl["verts_uvs"],
                         mano=pack, device=device)
    return hc, pack, tmpl


@torch.no_grad()
def render_frames(hc, p, S, focal, chunk=16, self_shadow=True):
    """Forward-only render of every frame of a parameter dict -> (sil (T,S,S), rgb (T,S,S,3), verts, joints)."""
    dev = hc.device
    T = p["pose"].shape[0]
    sil, rgb, vs, js = [], [], [], []
    for a in range(0, T, chunk):
        sl = slice(a, min(a + chunk, T))
        B = sl.stop - sl.start
        if hc.J == 55:
            full = SMPLXARM.assemble_full_pose(p["rot"][sl].to(dev), p["pose"][sl].to(dev), p["wrist_pose"][sl].to(dev))
            verts0, joints = Lbs.apply(hc, full, p["shape"].to(dev), p["trans"][sl].to(dev))
        else:
            pose48 = torch.cat((p["rot"][sl], p["pose"][sl]), 1).to(dev)
            verts0, joints = LbsMano.apply(hc, pose48, p["shape"].to(dev), p["trans"][sl].to(dev))
        verts, normals = MeshBuild.apply(hc, verts0, p["verts_disps"].to(dev))
        cR = camera.flip_xy_R(B, dev)
        cT = camera.cam_to_T(p["cam"][sl].to(dev), focal, S)
        light = p["light_positions"][0].to(dev)[None].expand(B, 3).contiguous()
        lR, lT = camera.light_camera(light, verts.mean(1))
        sil.append(Silhouette.apply(hc, verts, cR, cT, focal, S, 1e-7))
        rgb.append(Shade.apply(hc, verts, normals, cR, cT, lR, lT, light, torch.sigmoid(p["amb_ratio"].to(dev)),
                               p["texture"].to(dev), p["normal_map"].to(dev), focal, S, int(self_shadow)))
        vs.append(verts)
        js.append(joints)
    return torch.cat(sil), torch.cat(rgb), torch.cat(vs), torch.cat(js)


def make_scene(hc, T, S, seed=1, frame_offset=0, total_frames=None, frame_stride=1):
    """frame_offset / total_frames / frame_stride: this rank's shard of a longer sequence under static frame sharding —
    contiguous (stride 1: frames offset .. offset+T-1) or round-robin (stride R: frames offset, offset+R, ...)."""
    dev = hc.device
    focal = 1000.0 * S / 224.0  # utils/config_utils.py:11-12
    total = total_frames or T
    seq = synth.make_sequence(total, seed=seed, img_size=S)
    sl = slice(frame_offset, frame_offset + T * frame_stride, frame_stride)
    app = synth.make_appearance(seed=2)
    params = dict(pose=seq["pose"][sl], rot=seq["rot"][sl], trans=seq["trans"][sl], cam=seq["cam"][sl], shape=seq["shape"],
                  verts_disps=torch.zeros(hc.V, 1), texture=app["texture"], normal_map=app["normal_map"],
                  light_positions=app["light_position"][None].repeat(T, 1), amb_ratio=app["amb_ratio"])
    arm = hc.J == 55
    if arm:  # wrist-centred arm lying along -x (body_models.py:2342): smaller weak-perspective scale, shifted into frame
        params["wrist_pose"] = 0.1 * torch.sin(torch.arange(total)[:, None] * 0.3 + torch.tensor([0.0, 1.0, 2.0]))[sl].float()
        params["cam"] = params["cam"] * torch.tensor([0.35, 1.0, 1.0]) + torch.tensor([0.0, 0.15, 0.0])
    g = torch.Generator().manual_seed(100 + seed)
    tgt = dict(params)
    noise = 0.05 * torch.randn(total, 45, generator=g)
    tgt["pose"] = params["pose"] + noise[sl]
    tgt["cam"] = params["cam"] + torch.tensor([0.1, 0.003, -0.002])
    app2 = synth.make_appearance(seed=5)
    tgt["texture"], tgt["normal_map"] = app2["texture"], app2["normal_map"]
    sil, rgb, _, joints = render_frames(hc, tgt, S, focal)
    y_sil = (sil >= 0.5).float()
    frames = dict(y_true=rgb.clamp(0, 1).contiguous(), y_sil=y_sil, y_sil_eroded=synth.erode_mask(y_sil).contiguous(),
                  init_joints=(joints[:, :21] * 1000.0).contiguous())
    first = {k: (v[:1] if k in ("pose", "rot", "trans", "cam", "light_positions", "wrist_pose") else v) for k, v in params.items()}
    if frame_offset != 0:  # the ARAP reference is frame 0 of the WHOLE sequence on every rank
        first.update(pose=seq["pose"][:1], rot=seq["rot"][:1], trans=seq["trans"][:1])
        if arm:
            first["wrist_pose"] = 0.1 * torch.sin(torch.zeros(1, 1) + torch.tensor([0.0, 1.0, 2.0])).float()
    _, _, ref_verts, _ = render_frames(hc, first, S, focal)
    tmpl = synth.load_template("arm" if arm else "hand")
    offsets = dict(albedo=synth.make_reg_offsets(seed=3, std=1.0).to(dev).contiguous(),
                   normal=synth.make_reg_offsets(seed=4, std=2.0).to(dev).contiguous())
    return dict(params=params, frames=frames, ref_verts=ref_verts[:1].contiguous(),
                uv_mask=torch.tensor(tmpl["uv_mask"]), offsets=offsets, focal=focal, img_size=S)
