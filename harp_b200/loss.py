"""Drop-ins for the loss terms of `optimize_sequence.py:517-553`: `loss/arap.py`, `loss/kps_loss.py`,
`loss/texture_reg.py` and PyTorch3D's `mesh_laplacian_smoothing` / `mesh_normal_consistency`, each a thin call into the
fused value+gradient kernels (`harp_mesh_losses`, `harp_texture_reg`)."""
import torch

from . import ops


def _mesh_term(meshes, index, ref_verts=None):
    return ops.MeshTerm.apply(meshes.hc, index, meshes.verts_padded(), ref_verts, None, None, None)


def mesh_laplacian_smoothing(meshes, method="uniform"):
    if method != "uniform":
        raise NotImplementedError("only the uniform Laplacian is on HARP's path (optimize_sequence.py:536)")
    return _mesh_term(meshes, 0)


def mesh_normal_consistency(meshes):
    return _mesh_term(meshes, 1)


def arap_loss(meshes, ref_meshes, target_length: float = 0.0):
    """loss/arap.py:4-57: squared edge-length change vs. a one-mesh reference, in mm^2."""
    if len(ref_meshes) != 1:
        raise NotImplementedError("per-frame reference meshes are not used by HARP (optimize_sequence.py:429-435)")
    return _mesh_term(meshes, 2, ref_meshes.verts_padded().detach())


def kps_loss(gt_kps, pred_kps, use_arm=False, device="cuda"):
    """loss/kps_loss.py:4-17: gt in mm, pred in m, root-aligned, (mm/100)^2."""
    B, nj = pred_kps.shape[0], pred_kps.shape[1]
    from .structures import any_context
    # the kps term needs no topology, only a context on the same device whose joint count matches the layout of `pred_kps`
    # (21 for MANO; 22 for the arm, of which the kernel compares the first 21 — loss/kps_loss.py:7-8)
    hc = any_context(pred_kps.device, nj)
    if hc is None:
        raise ValueError(f"kps_loss: no registered context on {pred_kps.device} produces {nj} joints")
    dummy = torch.zeros(B, hc.V, 3, device=pred_kps.device)
    return ops.MeshTerm.apply(hc, 3, dummy, None, pred_kps, gt_kps.to(pred_kps.device), None)


def _draw_offsets(shape_hw, std):
    """The reference draws neighbour offsets from the CPU RNG on every call (loss/texture_reg.py:15,51); same call here,
    so a `torch.manual_seed` before it reproduces the reference's offsets exactly."""
    return torch.normal(mean=0, std=std, size=(*shape_hw, 2)).to(torch.int)


def albedo_reg(uv_texture, std=2.0, uv_mask=None):
    off = _draw_offsets(uv_texture.shape[-3:-1], std)
    return ops.TextureReg.apply(None, 0, uv_texture, off, uv_mask.to(uv_texture.device) if uv_mask is not None else None)


def normal_reg(normal_map, std=2.0, uv_mask=None):
    off = _draw_offsets(normal_map.shape[-3:-1], std)
    return ops.TextureReg.apply(None, 1, normal_map, off, uv_mask.to(normal_map.device) if uv_mask is not None else None)
