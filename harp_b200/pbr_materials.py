"""`renderer/pbr_materials.py:PBRMaterials` at the call boundary: a bag of material constants + the normal-map texture.

The TBN perturbation itself (`compute_tangent` 58-77, `apply_normal_map` 82-124) runs inside `shade_deferred_kernel`
(`harp_b200/csrc/shade.cu`); this class only carries what `render_image*` hands to the renderer."""
import torch


class PBRMaterials:
    def __init__(self, ambient_color=((1, 1, 1),), diffuse_color=((1, 1, 1),), specular_color=((1, 1, 1),), shininess=64,
                 normal_maps=None, device="cpu"):
        self.device = device
        as_t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=device)
        self.ambient_color, self.diffuse_color, self.specular_color = as_t(ambient_color), as_t(diffuse_color), as_t(specular_color)
        for name in ("ambient_color", "diffuse_color", "specular_color"):
            t = getattr(self, name)
            if t.shape[-1] != 3:
                raise ValueError("Expected %s to have shape (N, 3); got %r" % (name, t.shape))
        self.shininess = as_t(shininess).reshape(-1)
        self.normal_maps = normal_maps
        self.use_normal_map = normal_maps is not None
