"""Drop-in for the hot-path half of `utils/visualize.py`: prepare_mesh (16-88), prepare_materials (91-108),
render_image (258-285), render_image_with_RT (288-319).  Same arguments, same returns; tensors stay on the GPU."""
import torch

from . import camera, ops
from .pbr_materials import PBRMaterials
from .structures import Meshes, TexturesUV, TexturesVertex


_DEVICE_COPIES = {}


def _on_device(t, device):
    """Static tables (faces_uvs, verts_uvs) that callers keep on the host, as the reference's init_params does
    (optimize_sequence.py:222-223): one host->device copy per tensor version instead of one per call."""
    dev = torch.device(device)
    if t.device.type == dev.type and (dev.index is None or t.device.index == dev.index):
        return t
    key = (id(t), t._version, str(dev))
    hit = _DEVICE_COPIES.get(key)
    if hit is None or hit[0]() is not t:
        import weakref
        if len(_DEVICE_COPIES) > 64:
            _DEVICE_COPIES.clear()
        hit = (weakref.ref(t), t.to(dev))
        _DEVICE_COPIES[key] = hit
    return hit[1]


def prepare_mesh(params, fid, mano_layer, verts_textures, mesh_subdivider, global_pose, configs, device="cuda",
                 vis_normal=False, shared_texture=True, use_arm=False):
    """-> (hand_joints (B,21,3) m, hand_verts (B,V,3) m, faces (B,F,3), textures).
    `mesh_subdivider`: None for the coarse mesh, anything truthy for one SubdivideMeshes step (the subdivision tables live
    in `mano_layer.hc`)."""
    if configs.get("model_type", "harp") != "harp":
        raise NotImplementedError("html / nimble hand models need third-party code and assets (SURVEY.md section 2 row 23)")
    B = fid.shape[0]
    fid = fid.to(params["pose"].device)
    betas = params["shape"].to(device)
    disps = params.get("verts_disps")
    # utils/visualize.py:49-64: the mesh is subdivided only inside the `verts_disps is not None` branch
    hc = mano_layer.hc if (mesh_subdivider is not None and disps is not None) else mano_layer.hc_coarse
    if use_arm:  # visualize.py:36-40
        full_pose = mano_layer.assemble_full_pose(params["rot"][fid].to(device), params["pose"][fid].to(device),
                                                  params["wrist_pose"][fid].to(device))
        verts0, joints = ops.Lbs.apply(hc, full_pose, betas, params["trans"][fid].to(device))
    else:
        pose48 = torch.cat((params["rot"][fid], params["pose"][fid]), 1).to(device)
        verts0, joints = ops.LbsMano.apply(hc, pose48, betas, params["trans"][fid].to(device))  # metres (visualize.py:45-46)
    if disps is not None and disps.shape[1] != 1:
        raise NotImplementedError("free 3-D vertex offsets (VERT_DISPS_NORMALS=False) are not used by HARP's config")
    verts, _ = ops.MeshBuild.apply(hc, verts0, disps.to(device) if disps is not None else None)
    faces = hc.faces()[None].expand(B, -1, -1)
    if verts_textures:
        textures = TexturesVertex(params["verts_rgb"].repeat(B, 1, 1).float().to(device))
    else:
        uv_map = params["texture"][None, 0].to(device) if shared_texture else params["texture"].to(device)
        textures = TexturesUV(maps=uv_map, faces_uvs=_on_device(params["faces_uvs"], device), verts_uvs=_on_device(params["verts_uvs"], device))
    return joints, verts, faces, textures


def prepare_materials(params, batch_size, shared_texture=True, device="cuda"):
    """The normal map goes to the kernel RAW: the per-texel F.normalize of visualize.py:98 happens at fetch time."""
    if "normal_map" not in params:
        return {"normal_maps": None}
    nm = params["normal_map"][None, 0].to(device) if shared_texture else params["normal_map"].to(device)
    return {"normal_maps": TexturesUV(maps=nm, faces_uvs=_on_device(params["faces_uvs"], device),
                                      verts_uvs=_on_device(params["verts_uvs"], device))}


def render_image(mesh, cam, batch_size, renderer, img_size, focal_length, silhouette=False, device="cuda",
                 materials_properties=dict()):
    materials = PBRMaterials(device=device, shininess=0.0, **materials_properties)
    cam_t = camera.cam_to_T(cam.to(device), focal_length, img_size)
    R = camera.flip_xy_R(batch_size, device)
    out = renderer(mesh, principal_point=torch.tensor([(img_size / 2.0, img_size / 2.0)]), focal_length=focal_length, T=cam_t, R=R,
                   materials=materials, image_size=torch.tensor([(img_size, img_size)]))
    return out[:, :, :, 3] if silhouette else out[:, :, :, 0:3]


def render_image_with_RT(mesh, light_t, light_r, cam_t, cam_r, batch_size, renderer, img_size, focal_length, silhouette=False,
                         materials_properties=dict(), device="cuda"):
    materials = PBRMaterials(device=device, shininess=0.0, **materials_properties)
    out = renderer(mesh, principal_point=torch.tensor([(img_size / 2.0, img_size / 2.0)]), focal_length=focal_length,
                   T=light_t.to(device), R=light_r.to(device), cam_T=cam_t.to(device), cam_R=cam_r.to(device), materials=materials,
                   image_size=torch.tensor([(img_size, img_size)]))
    return out[:, :, :, 3] if silhouette else out[:, :, :, 0:3]


def _rotate_axis_angle(points, degrees, axis):
    """pytorch3d.transforms.RotateAxisAngle(degrees, axis).transform_points: row vectors times the transposed rotation."""
    a = torch.tensor(degrees / 180.0 * 3.141592653589793, dtype=points.dtype, device=points.device)
    c, s, o, z = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
    R = {"X": [o, z, z, z, c, -s, z, s, c], "Y": [c, z, s, z, o, z, -s, z, c], "Z": [c, -s, z, s, c, z, z, z, o]}[axis]
    M = torch.stack(R).view(3, 3).t()
    return points @ M


def render_360(params, fid, renderer, img_size, focal_length, mano_layer, configs, render_normal=False, verts_textures=True,
               mesh_subdivider=None, global_pose=False, global_betas=True, device="cuda", save_img_dir=None, use_arm=False):
    """Turntable of `utils/visualize.py:145-199`: 36 cumulative 10-degree turns about Y, then 36 about X, of the posed mesh around
    its centroid, each rendered with `renderer` (the phong or normal renderer of get_renderers).  Returns the 72 frames of the
    first sample as a uint8 tensor (72,S,S,3) on the device; with `save_img_dir` they are also written as the reference names
    them (`render_360[_normal]/%04d.jpg`, `h_%04d.jpg`; the animated GIF is left to the caller)."""
    joints, hand_verts, faces, textures = prepare_mesh(params, fid, mano_layer, verts_textures, mesh_subdivider, global_pose,
                                                       configs=configs, device=device, use_arm=use_arm)
    cam = params["cam"][fid.to(params["cam"].device)]
    B = fid.shape[0]
    materials_properties = prepare_materials(params, B, device=device)
    center = torch.mean(hand_verts, dim=1, keepdim=True)
    hand_verts = hand_verts.detach()
    frames = []
    for axis in ("Y", "X"):
        for _ in range(36):
            hand_verts = _rotate_axis_angle(hand_verts - center, 10.0, axis) + center
            img = render_image(Meshes(hand_verts, faces, textures), cam, B, renderer, img_size, focal_length,
                               materials_properties=materials_properties, device=device)
            frames.append((img.detach()[0].clip(0, 1) * 255).to(torch.uint8))  # astype(np.uint8): truncation
    frames = torch.stack(frames)
    if save_img_dir is not None:
        import os
        from PIL import Image
        out_dir = os.path.join(save_img_dir, "render_360_normal" if render_normal else "render_360")
        os.makedirs(out_dir, exist_ok=True)
        for i, f in enumerate(frames.cpu().numpy()):
            Image.fromarray(f).save(os.path.join(out_dir, ("%04d.jpg" % i) if i < 36 else ("h_%04d.jpg" % (i - 36))))
    return frames


def render_360_light(params, fid, hand_verts, faces, textures, img_size, focal_length, save_img_dir=None, device="cuda"):
    """Light sweep of `utils/visualize.py:199-228`: 40 renders of the posed mesh with the Phong renderer (no normal map: the
    reference passes no materials_properties here, 220) while the point light moves along z from -5 to 5 at x = y = 1.
    Returns the first sample's frames as uint8 (40,S,S,3) on the device; with `save_img_dir` they are also written as
    `render_360_light/%04d.jpg` (the animated GIF is left to the caller)."""
    from . import renderer_helper
    cam = params["cam"][fid.to(params["cam"].device)]
    B = fid.shape[0]
    mesh = Meshes(hand_verts.detach(), faces, textures)
    start, end, num = -5.0, 5.0, 40
    frames = []
    for idx in range(num):
        cur_z = start + (end - start) / num * idx
        light_positions = torch.Tensor((1.0, 1.0, cur_z)).repeat(B, 1)
        phong_renderer, _, _ = renderer_helper.get_renderers(image_size=img_size, light_posi=light_positions, silh_sigma=1e-7,
                                                             silh_gamma=1e-1, silh_faces_per_pixel=50, device=device)
        img = render_image(mesh, cam, B, phong_renderer, img_size, focal_length, device=device)
        frames.append((img.detach()[0].clip(0, 1) * 255).to(torch.uint8))
    frames = torch.stack(frames)
    if save_img_dir is not None:
        import os
        from PIL import Image
        out_dir = os.path.join(save_img_dir, "render_360_light")
        os.makedirs(out_dir, exist_ok=True)
        for i, f in enumerate(frames.cpu().numpy()):
            Image.fromarray(f).save(os.path.join(out_dir, "%04d.jpg" % i))
    return frames
