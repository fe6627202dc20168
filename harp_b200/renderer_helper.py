"""Drop-in for the factories of `renderer/renderer_helper.py`: same names, arguments and call convention; the returned
renderers launch the fused CUDA kernels instead of PyTorch3D's rasterise -> Fragments -> shader pipeline.

    phong, silhouette, normal = get_renderers(image_size, light_posi, silh_sigma, silh_gamma, silh_faces_per_pixel, device)
    shadow = get_shadow_renderers(image_size, light_posi, ..., amb_ratio, device)
    light_R, light_T, cam_R, cam_T = process_info_for_shadow(cam, light_positions, hand_verts_center, image_size, focal_length)
    image = renderer(meshes, principal_point=, focal_length=, T=, R=, [cam_T=, cam_R=,] materials=, image_size=)   # (B,S,S,4)
"""
import torch

from . import camera, ops


def _scalar(x):
    return float(x.reshape(-1)[0]) if torch.is_tensor(x) else float(x)


def _check_camera(kwargs, image_size):
    """The kernels implement PerspectiveCameras(in_ndc=False) with the principal point at the image centre
    (utils/visualize.py:263-279); anything else is outside the hot path."""
    pp = kwargs.get("principal_point")
    if pp is not None:
        pp = torch.as_tensor(pp, dtype=torch.float32).reshape(-1)
        if not torch.allclose(pp, torch.full_like(pp, image_size / 2.0)):
            raise ValueError("harp_b200 renderers need principal_point == image_size / 2")
    if "T" not in kwargs or "R" not in kwargs or "focal_length" not in kwargs:
        raise ValueError("Cameras must be specified in the forward pass: T=, R=, focal_length=")


class SilhouetteRenderer:
    """MeshRenderer(MeshRasterizer(K, blur=ln(1/1e-4-1) sigma), SoftSilhouetteShader) — renderer_helper.py:37-58."""

    def __init__(self, image_size, sigma, faces_per_pixel):
        self.image_size, self.sigma, self.faces_per_pixel = int(image_size), float(sigma), int(faces_per_pixel)

    def __call__(self, meshes, **kwargs):
        _check_camera(kwargs, self.image_size)
        alpha = ops.Silhouette.apply(meshes.hc, meshes.verts_padded(), kwargs["R"], kwargs["T"], _scalar(kwargs["focal_length"]),
                                     self.image_size, self.sigma, self.faces_per_pixel)
        return torch.cat([torch.ones_like(alpha)[..., None].expand(-1, -1, -1, 3), alpha[..., None]], dim=-1)


class ShadowRenderer:
    """MeshRendererShadow(MeshRasterizer(K=1), SoftPhongShaderShadow) — renderer_helper.py:331-451.
    Call with T/R = light camera, cam_T/cam_R = viewing camera, materials = PBRMaterials(normal_maps=TexturesUV)."""

    def __init__(self, image_size, light_posi, amb_ratio, self_shadow=True):
        self.image_size, self.light_posi, self.amb_ratio, self.self_shadow = int(image_size), light_posi, amb_ratio, self_shadow

    def __call__(self, meshes, **kwargs):
        _check_camera(kwargs, self.image_size)
        tex = meshes.textures
        mat = kwargs.get("materials")
        if tex is None or mat is None or not getattr(mat, "use_normal_map", False):
            raise ValueError("shadow renderer needs meshes.textures (TexturesUV) and materials with normal_maps")
        verts = meshes.verts_padded()
        B = verts.shape[0]
        light = torch.as_tensor(self.light_posi, dtype=torch.float32, device=verts.device).reshape(-1, 3).expand(B, 3)
        amb = torch.as_tensor(self.amb_ratio, dtype=torch.float32, device=verts.device).reshape(-1)[:1]
        cam_T, cam_R = kwargs.get("cam_T", kwargs["T"]), kwargs.get("cam_R", kwargs["R"])
        rgb = ops.Shade.apply(meshes.hc, verts, meshes.verts_normals_padded(), cam_R, cam_T, kwargs["R"], kwargs["T"], light, amb,
                              tex.maps_padded()[:1], mat.normal_maps.maps_padded()[:1], _scalar(kwargs["focal_length"]),
                              self.image_size, int(self.self_shadow))
        hit = ops.hit_mask(meshes.hc, B, self.image_size)  # alpha of a K=1 blend: 1 on covered pixels
        return torch.cat([rgb, hit], dim=-1)


class PhongRenderer:
    """MeshRenderer(MeshRasterizer(K=1), SoftPhongShaderPBR) — renderer_helper.py:60-83,106-190: point light at `light_posi`
    with ambient .5 / diffuse .4 / specular .1, shininess 0 (utils/visualize.py:259-263), no shadow pass."""
    LIGHTS = (0.5, 0.4, 0.1)

    def __init__(self, image_size, light_posi):
        self.image_size, self.light_posi = int(image_size), light_posi

    def __call__(self, meshes, **kwargs):
        _check_camera(kwargs, self.image_size)
        tex, mat = meshes.textures, kwargs.get("materials")
        if tex is None:
            raise ValueError("phong renderer needs meshes.textures (TexturesUV)")
        verts = meshes.verts_padded()
        B = verts.shape[0]
        if mat is not None and getattr(mat, "use_normal_map", False):
            nm = mat.normal_maps.maps_padded()[:1]
        else:  # materials without normal maps (render_360_light, utils/visualize.py:220): a constant +z map leaves the normals as they are
            t = tex.maps_padded()  # the kernel samples both maps on one (H, W) grid
            nm = verts.new_tensor([0.0, 0.0, 1.0]).expand(1, t.shape[-3], t.shape[-2], 3).contiguous()
        light = torch.as_tensor(self.light_posi, dtype=torch.float32, device=verts.device).reshape(-1, 3).expand(B, 3)
        amb = verts.new_zeros(1)
        rgb = ops.Shade.apply(meshes.hc, verts, meshes.verts_normals_padded(), kwargs["R"], kwargs["T"], kwargs["R"], kwargs["T"], light,
                              amb, tex.maps_padded()[:1], nm, _scalar(kwargs["focal_length"]),
                              self.image_size, 0, self.LIGHTS)
        hit = ops.hit_mask(meshes.hc, B, self.image_size)
        return torch.cat([rgb, hit], dim=-1)


class NormalRenderer:
    """MeshRenderer(MeshRasterizer(K=10), SoftPhongNormalShader) — renderer_helper.py:85-101,216-301.  Evaluation pass:
    forward only (the reference never differentiates through it)."""

    def __init__(self, image_size, faces_per_pixel=10):
        self.image_size, self.faces_per_pixel = int(image_size), int(faces_per_pixel)

    def __call__(self, meshes, **kwargs):
        _check_camera(kwargs, self.image_size)
        mat = kwargs.get("materials")
        nm = mat.normal_maps.maps_padded()[:1] if mat is not None and getattr(mat, "use_normal_map", False) else None
        verts = meshes.verts_padded()
        rgb = ops.render_normals(meshes.hc, verts, meshes.verts_normals_padded(), kwargs["R"], kwargs["T"],
                                 _scalar(kwargs["focal_length"]), self.image_size, nm, self.faces_per_pixel)
        hit = (rgb != 1.0).any(-1, keepdim=True).to(rgb.dtype)
        return torch.cat([rgb, hit], dim=-1)


def get_renderers(image_size, light_posi=((1.0, 1.0, -5.0),), silh_sigma=1e-7, silh_gamma=1e-1, silh_faces_per_pixel=50,
                  device="cuda"):
    """-> (phong_renderer, silhouette_renderer, normal_renderer) as renderer_helper.get_renderers (26-103)."""
    return (PhongRenderer(image_size, light_posi), SilhouetteRenderer(image_size, silh_sigma, silh_faces_per_pixel),
            NormalRenderer(image_size, 10))


def get_shadow_renderers(image_size, light_posi=((1.0, 1.0, -5.0),), silh_sigma=1e-7, silh_gamma=1e-1, silh_faces_per_pixel=50,
                         amb_ratio=0.6, device="cuda"):
    """renderer_helper.get_shadow_renderers (416-451): amb_ratio is the POST-sigmoid ambient (optimize_sequence.py:480)."""
    return ShadowRenderer(image_size, light_posi, amb_ratio)


def process_info_for_shadow(cam, light_positions, hand_verts_center, image_size, focal_length, device="cuda"):
    """renderer_helper.process_info_for_shadow (454-468) -> light_R, light_T, cam_R, cam_T."""
    dev = hand_verts_center.device
    cam_T = camera.cam_to_T(cam.to(dev), focal_length, image_size)
    cam_R = camera.flip_xy_R(cam.shape[0], dev)
    light_R, light_T = camera.light_camera(light_positions.to(dev), hand_verts_center)
    return light_R, light_T, cam_R, cam_T
