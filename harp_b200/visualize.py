"""Drop-in for the hot-path half of `utils/visualize.py`: prepare_mesh (16-88), prepare_materials (91-108),
render_image (258-285), render_image_with_RT (288-319).  Same arguments, same returns; tensors stay on the GPU."""
import torch

from . import camera, ops
from .pbr_materials import PBRMaterials
from .structures import Meshes, TexturesUV, TexturesVertex


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
    hc = mano_layer.hc if mesh_subdivider is not None else mano_layer.hc_coarse
    if use_arm:  # visualize.py:36-40
        full_pose = mano_layer.assemble_full_pose(params["rot"][fid].to(device), params["pose"][fid].to(device),
                                                  params["wrist_pose"][fid].to(device))
        verts0, joints = ops.Lbs.apply(hc, full_pose, betas, params["trans"][fid].to(device))
    else:
        pose48 = torch.cat((params["rot"][fid], params["pose"][fid]), 1).to(device)
        verts0, joints = ops.LbsMano.apply(hc, pose48, betas, params["trans"][fid].to(device))  # metres (visualize.py:45-46)
    disps = params.get("verts_disps")
    if disps is not None and disps.shape[1] != 1:
        raise NotImplementedError("free 3-D vertex offsets (VERT_DISPS_NORMALS=False) are not used by HARP's config")
    verts, _ = ops.MeshBuild.apply(hc, verts0, disps.to(device) if disps is not None else None)
    faces = hc.faces()[None].expand(B, -1, -1)
    if verts_textures:
        textures = TexturesVertex(params["verts_rgb"].repeat(B, 1, 1).float().to(device))
    else:
        uv_map = params["texture"][None, 0].to(device) if shared_texture else params["texture"].to(device)
        textures = TexturesUV(maps=uv_map, faces_uvs=params["faces_uvs"].to(device), verts_uvs=params["verts_uvs"].to(device))
    return joints, verts, faces, textures


def prepare_materials(params, batch_size, shared_texture=True, device="cuda"):
    """The normal map goes to the kernel RAW: the per-texel F.normalize of visualize.py:98 happens at fetch time."""
    if "normal_map" not in params:
        return {"normal_maps": None}
    nm = params["normal_map"][None, 0].to(device) if shared_texture else params["normal_map"].to(device)
    return {"normal_maps": TexturesUV(maps=nm, faces_uvs=params["faces_uvs"].to(device), verts_uvs=params["verts_uvs"].to(device))}


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
