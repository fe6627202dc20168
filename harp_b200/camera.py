"""Per-frame camera glue (tiny (B,3)/(B,3,3) tensors, plain torch so autograd covers them on any device).

Weak-perspective `cam = (s, tx, ty)` -> PyTorch3D-convention view transform `X_view = X_world R + T`
(`utils/visualize.py:266-270`), and the shadow pass's light camera (`renderer/renderer_helper.py:454-468`).
"""
import torch
import torch.nn.functional as F

SHADOW_SPHERE_RADIUS = 1.5  # renderer_helper.py:463


def cam_to_T(cam, focal_length, img_size):
    """(B,3) weak-perspective (s, tx, ty) -> (B,3) translation; depth = 2 f / (S s + 1e-9)."""
    tz = 2 * focal_length / (img_size * cam[:, 0] + 1e-9)
    return torch.stack([-cam[:, 1], -cam[:, 2], tz], dim=1)


_CONST = {}


def _const(name, values, device, dtype):
    """Small constants cached per device so that no host->device copy happens inside a hot (graph-captured) step."""
    key = (name, str(device), dtype)
    if key not in _CONST:
        _CONST[key] = torch.tensor(values, device=device, dtype=dtype)
    return _CONST[key]


def flip_xy_R(batch_size, device, dtype=torch.float32):
    """OpenCV -> PyTorch3D axis flip, diag(-1,-1,1), repeated per frame."""
    R = _const("flip", [[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, 1.0]], device, dtype)
    return R.expand(batch_size, 3, 3).contiguous()


def look_at_rotation(camera_position, at, up=(0.0, 1.0, 0.0)):
    """PyTorch3D look_at_rotation: columns of R are the camera x, y, z axes in world space."""
    up_t = _const("up" + repr(tuple(up)), list(up), camera_position.device, camera_position.dtype).expand_as(camera_position)
    z = F.normalize(at - camera_position, eps=1e-5)
    x = F.normalize(torch.cross(up_t, z, dim=1), eps=1e-5)
    y = F.normalize(torch.cross(z, x, dim=1), eps=1e-5)
    degenerate = (x.abs() <= 5e-3).all(dim=1, keepdim=True)
    x = torch.where(degenerate, F.normalize(torch.cross(y, z, dim=1), eps=1e-5), x)
    return torch.stack((x, y, z), dim=2)


def light_camera(light_positions, centroid):
    """Light camera on a sphere of radius 1.5 around the mesh centroid, looking at it -> (R (B,3,3), T (B,3))."""
    offset = light_positions - centroid
    eye = centroid + offset * (SHADOW_SPHERE_RADIUS / torch.linalg.norm(offset, dim=1, keepdim=True))
    R = look_at_rotation(eye, centroid)
    T = -torch.bmm(R.transpose(1, 2), eye[:, :, None])[:, :, 0]
    return R, T
