"""ORACLE (test infrastructure, never shipped): minimal restatement of the PyTorch3D v0.6.2 API that HARP imports.

PARITY UNPINNED at this layer (see `_raster.py` header): PyTorch3D is un-vendored and absent here, so these classes
restate its published behaviour, exposing exactly the symbols the reference uses:
`renderer/renderer_helper.py:4-23,471`, `renderer/pbr_materials.py:3-4`, `utils/visualize.py:6-11`,
`loss/arap.py:25-41`, `optimize_sequence.py:21-24`.  With this package first on sys.path the reference's own
`renderer_helper.py`, `pbr_materials.py`, `visualize.py`, `arap.py` run unmodified on CPU (tests/golden/make_golden.py).
"""
import math
from collections import namedtuple
from typing import NamedTuple, Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _raster

Device = Union[str, torch.device]


# ------------------------------------------------------------------------------------------------- utils
def _to_tensor(x, dtype, device):
    if torch.is_tensor(x):
        t = x.to(device=device)
        if t.dtype != dtype and t.is_floating_point():
            t = t.to(dtype)
        return t
    return torch.tensor(x, dtype=dtype, device=device)


class TensorProperties(nn.Module):
    """renderer/utils.py: broadcasting bag of per-batch tensors."""

    def __init__(self, dtype=torch.float32, device: Device = "cpu", **kwargs):
        super().__init__()
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self._N = 0
        if kwargs:
            vals = {}
            n = 1
            for k, v in kwargs.items():
                if v is None or isinstance(v, (bool, str)):
                    setattr(self, k, v)
                    continue
                t = _to_tensor(v, dtype, self.device)
                if t.dim() == 0:
                    t = t.view(1)
                vals[k] = t
                n = max(n, t.shape[0])
            for k, t in vals.items():
                if t.shape[0] != n:
                    if t.shape[0] != 1:
                        raise ValueError("Got non-broadcastable sizes")
                    t = t.expand(n, *t.shape[1:])
                setattr(self, k, t)
            self._N = n

    def __len__(self):
        return self._N

    def isempty(self):
        return self._N == 0

    def to(self, device="cpu"):
        self.device = torch.device(device)
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v) and v.device != self.device:
                setattr(self, k, v.to(self.device))
        return self

    def clone(self, other=None):
        if other is None:
            other = self.__class__(device=self.device)
        for k, v in self.__dict__.items():
            if k.startswith("_") and k != "_N":
                continue
            if torch.is_tensor(v):
                setattr(other, k, v.clone())
            elif k in ("_N", "device") or isinstance(v, (int, float, bool, str, type(None))):
                setattr(other, k, v)
            else:
                setattr(other, k, v)
        return other


def convert_to_tensors_and_broadcast(*args, dtype=torch.float32, device="cpu"):
    ts = [_to_tensor(a, dtype, device) for a in args]
    ts = [t.view(1, *t.shape) if t.dim() < 2 and False else t for t in ts]
    n = max(t.shape[0] for t in ts)
    out = []
    for t in ts:
        if t.shape[0] != n:
            if t.shape[0] != 1:
                raise ValueError("Got non-broadcastable sizes")
            t = t.expand(n, *t.shape[1:])
        out.append(t)
    return out


# ------------------------------------------------------------------------------------------------- transforms
class Transform3d:
    """transforms/transform3d.py, row-vector convention: p' = [p,1] @ M, then homogeneous divide."""

    def __init__(self, matrix=None, device="cpu", dtype=torch.float32):
        if matrix is None:
            matrix = torch.eye(4, dtype=dtype, device=device).view(1, 4, 4)
        self._matrix = matrix.view(-1, 4, 4)
        self._transforms = []

    def compose(self, *others):
        out = Transform3d(matrix=self._matrix)
        out._transforms = self._transforms + list(others)
        return out

    def get_matrix(self):
        m = self._matrix
        for o in self._transforms:
            om = o.get_matrix()
            if m.shape[0] != om.shape[0]:
                if m.shape[0] == 1:
                    m = m.expand(om.shape[0], -1, -1)
                elif om.shape[0] == 1:
                    om = om.expand(m.shape[0], -1, -1)
            m = m.bmm(om)
        return m

    def inverse(self):
        return Transform3d(matrix=torch.inverse(self.get_matrix()))

    def transform_points(self, points, eps=None):
        pb = points.clone()
        if pb.dim() == 2:
            pb = pb[None]
        ones = torch.ones(*pb.shape[:2], 1, dtype=pb.dtype, device=pb.device)
        pb = torch.cat([pb, ones], dim=2)
        m = self.get_matrix().to(pb.dtype)
        if m.shape[0] != pb.shape[0]:
            if m.shape[0] == 1:
                m = m.expand(pb.shape[0], -1, -1)
            else:
                pb = pb.expand(m.shape[0], -1, -1)
        out = pb.bmm(m)
        den = out[..., 3:]
        if eps is not None:
            sign = den.sign() + (den == 0.0).type_as(den)
            den = sign * torch.clamp(den.abs(), eps)
        out = out[..., :3] / den
        if out.shape[0] == 1 and points.dim() == 2:
            out = out.reshape(points.shape)
        return out


class RotateAxisAngle(Transform3d):  # utils/visualize.py:11 (turntables only; not on the hot path)
    def __init__(self, angle, axis="X", degrees=True, dtype=torch.float32, device="cpu"):
        a = torch.as_tensor(angle, dtype=dtype, device=device).view(-1)
        if degrees:
            a = a / 180.0 * math.pi
        c, s, o, z = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
        if axis == "X":
            R = [o, z, z, z, c, -s, z, s, c]
        elif axis == "Y":
            R = [c, z, s, z, o, z, -s, z, c]
        else:
            R = [c, -s, z, s, c, z, z, z, o]
        R = torch.stack(R, -1).view(-1, 3, 3).transpose(1, 2)
        m = torch.eye(4, dtype=dtype, device=device).repeat(a.shape[0], 1, 1)
        m[:, :3, :3] = R
        super().__init__(matrix=m)


def look_at_rotation(camera_position, at=((0, 0, 0),), up=((0, 1, 0),), device: Device = "cpu"):
    """renderer/cameras.py look_at_rotation."""
    dtype = camera_position.dtype if torch.is_tensor(camera_position) else torch.float32
    camera_position, at, up = convert_to_tensors_and_broadcast(camera_position, at, up, dtype=dtype, device=device)
    z_axis = F.normalize(at - camera_position, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    is_close = torch.isclose(x_axis, torch.tensor(0.0, dtype=dtype), atol=5e-3).all(dim=1, keepdim=True)
    if is_close.any():
        replacement = F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5)
        x_axis = torch.where(is_close, replacement, x_axis)
    R = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    return R.transpose(1, 2)


# ------------------------------------------------------------------------------------------------- cameras
def get_ndc_to_screen_transform(cameras, with_xyflip=False, image_size=None):
    image_size = _to_tensor(image_size, torch.float32, cameras.device).view(-1, 2)
    n = max(cameras._N, image_size.shape[0])
    dt = cameras._dtype()
    K = torch.zeros((n, 4, 4), dtype=dt, device=cameras.device)
    image_size = image_size.to(dt).expand(n, 2)
    height, width = image_size.unbind(1)
    scale = (image_size.min(dim=1).values - 0.0) / 2.0
    K[:, 0, 0] = scale
    K[:, 1, 1] = scale
    K[:, 0, 3] = -1.0 * (width - 0.0) / 2.0
    K[:, 1, 3] = -1.0 * (height - 0.0) / 2.0
    K[:, 2, 2] = 1.0
    K[:, 3, 3] = 1.0
    t = Transform3d(matrix=K.transpose(1, 2).contiguous())
    if with_xyflip:
        f = torch.eye(4, dtype=dt, device=cameras.device)
        f[0, 0] = -1.0
        f[1, 1] = -1.0
        t = t.compose(Transform3d(matrix=f.view(1, 4, 4)))
    return t


class PerspectiveCameras(TensorProperties):
    """renderer/cameras.py PerspectiveCameras, screen-space (in_ndc=False) path only."""

    def __init__(self, focal_length=1.0, principal_point=((0.0, 0.0),), R=None, T=None, K=None,
                 device: Device = "cpu", in_ndc: bool = True, image_size=None):
        if R is None:
            R = torch.eye(3)[None]
        if T is None:
            T = torch.zeros(1, 3)
        kw = dict(focal_length=focal_length, principal_point=principal_point, R=R, T=T)
        if image_size is not None:
            kw["image_size"] = image_size
        super().__init__(device=device, **kw)
        self._in_ndc = in_ndc
        if image_size is None:
            self.image_size = None

    def _dtype(self):
        return self.T.dtype if torch.is_tensor(self.T) else torch.float32

    def in_ndc(self):
        return self._in_ndc

    def is_perspective(self):
        return True

    def get_znear(self):
        return getattr(self, "znear", None)

    def get_image_size(self):
        return self.image_size

    def clone(self):
        other = PerspectiveCameras(device=self.device, in_ndc=self._in_ndc)
        for k in ("focal_length", "principal_point", "R", "T", "image_size"):
            v = getattr(self, k, None)
            setattr(other, k, v.clone() if torch.is_tensor(v) else v)
        other._N = self._N
        return other

    def get_world_to_view_transform(self, **kwargs):
        # NB: stores kwargs R/T on the camera (side effect relied on at renderer_helper.py:349-352)
        R = kwargs.get("R", self.R)
        T = kwargs.get("T", self.T)
        self.R = R
        self.T = T
        if T.shape[0] != R.shape[0]:
            n = max(T.shape[0], R.shape[0])
            R = R.expand(n, 3, 3)
            T = T.expand(n, 3)
        n = R.shape[0]
        Rm = torch.zeros(n, 4, 4, dtype=R.dtype, device=R.device)
        Rm[:, :3, :3] = R
        Rm[:, 3, 3] = 1.0
        Tm = torch.eye(4, dtype=R.dtype, device=R.device).repeat(n, 1, 1)
        Tm[:, 3, :3] = T.to(R.dtype)
        return Transform3d(matrix=Rm).compose(Transform3d(matrix=Tm))

    def get_camera_center(self, **kwargs):
        w2v = self.get_world_to_view_transform(**kwargs)
        P = w2v.inverse().get_matrix()
        return P[:, 3, :3]

    def get_projection_transform(self, **kwargs):
        fl = kwargs.get("focal_length", self.focal_length)
        pp = kwargs.get("principal_point", self.principal_point)
        dt = self._dtype()
        fl = _to_tensor(fl, dt, self.device).to(dt)
        pp = _to_tensor(pp, dt, self.device).to(dt).view(-1, 2)
        if fl.dim() in (0, 1) or fl.shape[1] == 1:
            fx = fy = fl.view(-1)
        else:
            fx, fy = fl.unbind(1)
        n = max(self._N, pp.shape[0], fx.shape[0])
        K = torch.zeros((n, 4, 4), dtype=dt, device=self.device)
        K[:, 0, 0] = fx
        K[:, 1, 1] = fy
        K[:, 0, 2] = pp[:, 0]
        K[:, 1, 2] = pp[:, 1]
        K[:, 3, 2] = 1.0
        K[:, 2, 3] = 1.0
        return Transform3d(matrix=K.transpose(1, 2).contiguous())

    def get_ndc_camera_transform(self, **kwargs):
        if self._in_ndc:
            return Transform3d(device=self.device)
        dt = self._dtype()
        pp = _to_tensor(kwargs.get("principal_point", self.principal_point), dt, self.device).to(dt).view(-1, 2)
        n = max(self._N, pp.shape[0])
        fix = torch.eye(4, dtype=dt, device=self.device).repeat(n, 1, 1)
        fix[:, :2, 3] = -2.0 * pp
        fix_t = Transform3d(matrix=fix.transpose(1, 2).contiguous())
        image_size = kwargs.get("image_size", self.get_image_size())
        s2n = get_ndc_to_screen_transform(self, with_xyflip=False, image_size=image_size).inverse()
        return fix_t.compose(s2n)

    def get_full_projection_transform(self, **kwargs):
        return self.get_world_to_view_transform(**kwargs).compose(self.get_projection_transform(**kwargs))

    def transform_points_ndc(self, points, eps=None, **kwargs):
        t = self.get_full_projection_transform(**kwargs).compose(self.get_ndc_camera_transform(**kwargs))
        return t.transform_points(points, eps=eps)

    def transform_points_screen(self, points, eps=None, **kwargs):
        ndc = self.transform_points_ndc(points, eps=eps, **kwargs)
        image_size = kwargs.get("image_size", self.get_image_size())
        return get_ndc_to_screen_transform(self, with_xyflip=True, image_size=image_size).transform_points(ndc, eps=eps)


# ------------------------------------------------------------------------------------------------- structures
class Meshes:
    """structures/meshes.py for homogeneous batches."""

    def __init__(self, verts=None, faces=None, textures=None):
        if isinstance(verts, (list, tuple)):
            verts = torch.stack(list(verts), 0)
        if isinstance(faces, (list, tuple)):
            faces = torch.stack(list(faces), 0)
        if verts.dim() == 2:
            verts = verts[None]
        if faces.dim() == 2:
            faces = faces[None]
        self._verts = verts
        self._faces = faces.long()
        self.textures = textures
        self.device = verts.device
        self._N, self._V = verts.shape[0], verts.shape[1]
        self._F = faces.shape[1]
        self._cache = {}

    def __len__(self):
        return self._N

    def isempty(self):
        return self._N == 0 or self._V == 0

    def __getitem__(self, i):
        if isinstance(i, int):
            i = slice(i, i + 1)
        tex = self.textures[i] if self.textures is not None and hasattr(self.textures, "__getitem__") else None
        return Meshes(self._verts[i], self._faces[i], tex)

    def to(self, device):
        return Meshes(self._verts.to(device), self._faces.to(device), self.textures)

    def verts_padded(self):
        return self._verts

    def faces_padded(self):
        return self._faces

    def verts_list(self):
        return list(self._verts.unbind(0))

    def faces_list(self):
        return list(self._faces.unbind(0))

    def verts_packed(self):
        return self._verts.reshape(-1, 3)

    def faces_packed(self):
        if "fp" not in self._cache:
            off = (torch.arange(self._N, device=self.device) * self._V).view(-1, 1, 1)
            self._cache["fp"] = (self._faces + off).reshape(-1, 3)
        return self._cache["fp"]

    def num_verts_per_mesh(self):
        return torch.full((self._N,), self._V, dtype=torch.long, device=self.device)

    def num_faces_per_mesh(self):
        return torch.full((self._N,), self._F, dtype=torch.long, device=self.device)

    def verts_packed_to_mesh_idx(self):
        return torch.arange(self._N, device=self.device).repeat_interleave(self._V)

    def faces_packed_to_mesh_idx(self):
        return torch.arange(self._N, device=self.device).repeat_interleave(self._F)

    def _compute_edges(self):
        if "edges" in self._cache:
            return
        faces = self.faces_packed()
        Fp = faces.shape[0]
        v0, v1, v2 = faces.chunk(3, dim=1)
        e01 = torch.cat([v0, v1], dim=1)
        e12 = torch.cat([v1, v2], dim=1)
        e20 = torch.cat([v2, v0], dim=1)
        edges = torch.cat([e12, e20, e01], dim=0)
        edge_to_mesh = torch.cat([self.faces_packed_to_mesh_idx()] * 3, 0)
        edges, _ = edges.sort(dim=1)
        V = self._N * self._V
        h = V * edges[:, 0] + edges[:, 1]
        u, inv = torch.unique(h, return_inverse=True)
        sorted_hash, sort_idx = torch.sort(h, dim=0)
        mask = torch.ones(h.shape[0], dtype=torch.bool, device=self.device)
        mask[1:] = sorted_hash[1:] != sorted_hash[:-1]
        uidx = sort_idx[mask]
        self._cache["edges"] = torch.stack([u // V, u % V], dim=1)
        self._cache["e2m"] = edge_to_mesh[uidx]
        self._cache["f2e"] = inv.reshape(3, Fp).t()
        self._cache["nE"] = torch.bincount(self._cache["e2m"], minlength=self._N)

    def edges_packed(self):
        self._compute_edges()
        return self._cache["edges"]

    def edges_packed_to_mesh_idx(self):
        self._compute_edges()
        return self._cache["e2m"]

    def faces_packed_to_edges_packed(self):
        self._compute_edges()
        return self._cache["f2e"]

    def num_edges_per_mesh(self):
        self._compute_edges()
        return self._cache["nE"]

    def verts_normals_packed(self):
        if "vn" not in self._cache:
            faces = self.faces_packed()
            verts = self.verts_packed()
            vn = torch.zeros_like(verts)
            vf = verts[faces]
            vn = vn.index_add(0, faces[:, 1], torch.cross(vf[:, 2] - vf[:, 1], vf[:, 0] - vf[:, 1], dim=1))
            vn = vn.index_add(0, faces[:, 2], torch.cross(vf[:, 0] - vf[:, 2], vf[:, 1] - vf[:, 2], dim=1))
            vn = vn.index_add(0, faces[:, 0], torch.cross(vf[:, 1] - vf[:, 0], vf[:, 2] - vf[:, 0], dim=1))
            self._cache["vn"] = F.normalize(vn, eps=1e-6, dim=1)
        return self._cache["vn"]

    def verts_normals_padded(self):
        return self.verts_normals_packed().view(self._N, self._V, 3)

    def laplacian_packed(self):
        if "L" not in self._cache:
            with torch.no_grad():
                verts, edges = self.verts_packed(), self.edges_packed()
                V = verts.shape[0]
                e0, e1 = edges.unbind(1)
                idx = torch.cat([torch.stack([e0, e1], 1), torch.stack([e1, e0], 1)], 0).t()
                ones = torch.ones(idx.shape[1], dtype=verts.dtype, device=self.device)
                A = torch.sparse_coo_tensor(idx, ones, (V, V))
                deg = torch.sparse.sum(A, dim=1).to_dense()
                d0 = deg[e0]
                d0 = torch.where(d0 > 0.0, 1.0 / d0, d0)
                d1 = deg[e1]
                d1 = torch.where(d1 > 0.0, 1.0 / d1, d1)
                L = torch.sparse_coo_tensor(idx, torch.cat([d0, d1]), (V, V))
                di = torch.arange(V, device=self.device)
                L = L - torch.sparse_coo_tensor(torch.stack([di, di], 0), torch.ones(V, dtype=verts.dtype), (V, V))
                self._cache["L"] = L.coalesce()
        return self._cache["L"]

    def update_padded(self, new_verts_padded):
        m = Meshes(new_verts_padded, self._faces, self.textures)
        for k in ("fp", "edges", "e2m", "f2e", "nE"):
            if k in self._cache:
                m._cache[k] = self._cache[k]
        return m

    def extend(self, N):
        return Meshes(self._verts.repeat_interleave(N, 0), self._faces.repeat_interleave(N, 0), self.textures)

    def sample_textures(self, fragments):
        return self.textures.sample_textures(fragments, faces_packed=self.faces_packed())


# ------------------------------------------------------------------------------------------------- ops
def interpolate_face_attributes(pix_to_face, barycentric_coords, face_attributes):
    """ops/interp_face_attrs.py."""
    N, H, W, K = pix_to_face.shape
    mask = pix_to_face < 0
    idx = pix_to_face.clamp(min=0).view(-1)
    attrs = face_attributes[idx]  # (P,3,D)
    out = (barycentric_coords.reshape(-1, 3, 1) * attrs).sum(1)
    out = out.view(N, H, W, K, -1)
    return out.masked_fill(mask[..., None], 0.0)


class SubdivideMeshes(nn.Module):
    """ops/subdivide_meshes.py."""

    def __init__(self, meshes=None):
        super().__init__()
        self.precomputed = False
        if meshes is not None:
            m0 = meshes[0]
            V = m0.verts_packed().shape[0]
            with torch.no_grad():
                fp = m0.faces_packed()
                fe = m0.faces_packed_to_edges_packed() + V
                f0 = torch.stack([fp[:, 0], fe[:, 2], fe[:, 1]], dim=1)
                f1 = torch.stack([fp[:, 1], fe[:, 0], fe[:, 2]], dim=1)
                f2 = torch.stack([fp[:, 2], fe[:, 1], fe[:, 0]], dim=1)
                f3 = fe
                self.register_buffer("_subdivided_faces", torch.cat([f0, f1, f2, f3], dim=0))
                self.register_buffer("_edges", m0.edges_packed())
            self.precomputed = True

    def forward(self, meshes, feats=None):
        verts = meshes.verts_padded()
        N = verts.shape[0]
        new_verts = verts[:, self._edges].mean(dim=2)
        new_verts = torch.cat([verts, new_verts], dim=1)
        new_faces = self._subdivided_faces.view(1, -1, 3).expand(N, -1, -1)
        return Meshes(verts=new_verts, faces=new_faces)


# ------------------------------------------------------------------------------------------------- textures
class TexturesVertex:
    def __init__(self, verts_features):
        if isinstance(verts_features, (list, tuple)):
            verts_features = torch.stack(list(verts_features), 0)
        self._feats = verts_features

    def to(self, device):
        return TexturesVertex(self._feats.to(device))

    def verts_features_padded(self):
        return self._feats

    def sample_textures(self, fragments, faces_packed=None):
        feats = self._feats.reshape(-1, self._feats.shape[-1])
        return interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords, feats[faces_packed])


class TexturesUV:
    """renderer/mesh/textures.py TexturesUV.sample_textures."""

    def __init__(self, maps, faces_uvs, verts_uvs, padding_mode="border", align_corners=True, sampling_mode="bilinear"):
        self._maps, self._faces_uvs, self._verts_uvs = maps, faces_uvs, verts_uvs
        self.padding_mode, self.align_corners, self.sampling_mode = padding_mode, align_corners, sampling_mode

    def to(self, device):
        return TexturesUV(self._maps.to(device), self._faces_uvs.to(device), self._verts_uvs.to(device),
                          self.padding_mode, self.align_corners, self.sampling_mode)

    def __getitem__(self, i):
        return TexturesUV(self._maps[i], self._faces_uvs[i], self._verts_uvs[i])

    def maps_padded(self):
        return self._maps

    def sample_textures(self, fragments, **kwargs):
        N = self._maps.shape[0]
        faces_verts_uvs = torch.cat([self._verts_uvs[n][self._faces_uvs[n].long()] for n in range(N)], 0)
        texture_maps = self._maps
        pixel_uvs = interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords, faces_verts_uvs)
        N, H_out, W_out, K = fragments.pix_to_face.shape
        N, H_in, W_in, C = texture_maps.shape
        pixel_uvs = pixel_uvs.permute(0, 3, 1, 2, 4).reshape(N * K, H_out, W_out, 2)
        texture_maps = (texture_maps.permute(0, 3, 1, 2)[None, ...].expand(K, -1, -1, -1, -1)
                        .transpose(0, 1).reshape(N * K, C, H_in, W_in))
        # v0.6.2: flip the map along H and map uv to [-1, 1]
        texture_maps = torch.flip(texture_maps, [2])
        pixel_uvs = pixel_uvs * 2.0 - 1.0
        texels = F.grid_sample(texture_maps, pixel_uvs, mode=self.sampling_mode,
                               align_corners=self.align_corners, padding_mode=self.padding_mode)
        return texels.reshape(N, K, C, H_out, W_out).permute(0, 3, 4, 1, 2)


# ------------------------------------------------------------------------------------------------- lighting / materials
class Materials(TensorProperties):
    def __init__(self, ambient_color=((1, 1, 1),), diffuse_color=((1, 1, 1),), specular_color=((1, 1, 1),),
                 shininess=64, device: Device = "cpu"):
        super().__init__(device=device, diffuse_color=diffuse_color, ambient_color=ambient_color,
                         specular_color=specular_color, shininess=shininess)

    def clone(self):
        return super().clone(Materials(device=self.device))


def _light_diffuse(normals, color, direction):
    points_dims = normals.shape[1:-1]
    expand_dims = (-1,) + (1,) * len(points_dims) + (3,)
    if direction.shape != normals.shape:
        direction = direction.view(expand_dims) if direction.dim() == 2 else direction
    if color.shape != normals.shape:
        color = color.view(expand_dims)
    normals = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    direction = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    angle = F.relu(torch.sum(normals * direction, dim=-1))
    return color * angle[..., None]


def _light_specular(points, normals, direction, color, camera_position, shininess):
    points_dims = points.shape[1:-1]
    expand_dims = (-1,) + (1,) * len(points_dims)
    if direction.shape != normals.shape:
        direction = direction.view(expand_dims + (3,)) if direction.dim() == 2 else direction
    if color.shape != normals.shape:
        color = color.view(expand_dims + (3,))
    if camera_position.shape != normals.shape:
        camera_position = camera_position.view(expand_dims + (3,))
    if shininess.shape != normals.shape:
        shininess = shininess.view(expand_dims)
    normals = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    direction = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    cos_angle = torch.sum(normals * direction, dim=-1)
    mask = (cos_angle > 0).to(normals.dtype)
    view_direction = F.normalize(camera_position - points, p=2, dim=-1, eps=1e-6)
    reflect_direction = -direction + 2 * (cos_angle[..., None] * normals)
    alpha = F.relu(torch.sum(view_direction * reflect_direction, dim=-1)) * mask
    return color * torch.pow(alpha, shininess)[..., None]


class PointLights(TensorProperties):
    """renderer/lighting.py PointLights."""

    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), location=((0, 1, 0),), device: Device = "cpu"):
        dtype = location.dtype if torch.is_tensor(location) else torch.float32
        super().__init__(dtype=dtype, device=device, ambient_color=ambient_color, diffuse_color=diffuse_color,
                         specular_color=specular_color, location=location)

    def clone(self):
        return super().clone(PointLights(device=self.device))

    def reshape_location(self, points):
        if self.location.ndim == points.ndim:
            return self.location
        return self.location[:, None, None, None, :]

    def diffuse(self, normals, points):
        direction = self.reshape_location(points) - points
        return _light_diffuse(normals, self.diffuse_color, direction)

    def specular(self, normals, points, camera_position, shininess):
        direction = self.reshape_location(points) - points
        return _light_specular(points, normals, direction, self.specular_color, camera_position, shininess)


def _apply_lighting(points, normals, lights, cameras, materials):
    """renderer/mesh/shading.py _apply_lighting."""
    light_diffuse = lights.diffuse(normals=normals, points=points)
    light_specular = lights.specular(normals=normals, points=points,
                                     camera_position=cameras.get_camera_center(), shininess=materials.shininess)
    ambient_color = materials.ambient_color * lights.ambient_color
    diffuse_color = materials.diffuse_color.view(-1, 1, 1, 1, 3) * light_diffuse
    specular_color = materials.specular_color.view(-1, 1, 1, 1, 3) * light_specular
    if ambient_color.ndim != diffuse_color.ndim:
        ambient_color = ambient_color[:, None, None, None, :]
    return ambient_color, diffuse_color, specular_color


# ------------------------------------------------------------------------------------------------- blending
class BlendParams(NamedTuple):
    sigma: float = 1e-4
    gamma: float = 1e-4
    background_color: Union[torch.Tensor, Sequence[float]] = (1.0, 1.0, 1.0)


def sigmoid_alpha_blend(colors, fragments, blend_params):
    """renderer/blending.py."""
    N, H, W, K = fragments.pix_to_face.shape
    pixel_colors = torch.ones((N, H, W, 4), dtype=colors.dtype, device=colors.device)
    mask = fragments.pix_to_face >= 0
    prob = torch.sigmoid(-fragments.dists / blend_params.sigma) * mask
    alpha = torch.prod((1.0 - prob), dim=-1)
    pixel_colors = torch.cat([colors[..., 0, :3], (1.0 - alpha)[..., None]], dim=-1)
    return pixel_colors


def softmax_rgb_blend(colors, fragments, blend_params, znear=1.0, zfar=100):
    N, H, W, K = fragments.pix_to_face.shape
    device = fragments.pix_to_face.device
    background = blend_params.background_color
    if not torch.is_tensor(background):
        background = torch.tensor(background, dtype=colors.dtype, device=device)
    else:
        background = background.to(device=device, dtype=colors.dtype)
    eps = 1e-10
    mask = fragments.pix_to_face >= 0
    prob_map = torch.sigmoid(-fragments.dists / blend_params.sigma) * mask
    alpha = torch.prod((1.0 - prob_map), dim=-1)
    z_inv = (zfar - fragments.zbuf) / (zfar - znear) * mask
    z_inv_max = torch.max(z_inv, dim=-1).values[..., None].clamp(min=eps)
    weights_num = prob_map * torch.exp((z_inv - z_inv_max) / blend_params.gamma)
    delta = torch.exp((eps - z_inv_max) / blend_params.gamma).clamp(min=eps)
    denom = weights_num.sum(dim=-1)[..., None] + delta
    weighted_colors = (weights_num[..., None] * colors).sum(dim=-2)
    weighted_background = delta * background
    rgb = (weighted_colors + weighted_background) / denom
    return torch.cat([rgb, (1.0 - alpha)[..., None]], dim=-1)


# ------------------------------------------------------------------------------------------------- rasteriser / renderer
class Fragments(NamedTuple):  # positional order relied on at renderer_helper.py:354-364,391,397
    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: torch.Tensor
    dists: torch.Tensor


class RasterizationSettings:
    def __init__(self, image_size=256, blur_radius=0.0, faces_per_pixel=1, bin_size=None, max_faces_per_bin=None,
                 perspective_correct=None, clip_barycentric_coords=None, cull_backfaces=False, z_clip_value=None,
                 cull_to_frustum=False):
        self.image_size, self.blur_radius, self.faces_per_pixel = image_size, float(blur_radius), faces_per_pixel
        self.perspective_correct, self.clip_barycentric_coords = perspective_correct, clip_barycentric_coords
        self.cull_backfaces, self.z_clip_value = cull_backfaces, z_clip_value


class MeshRasterizer(nn.Module):
    """renderer/mesh/rasterizer.py."""
    last_overflow = False

    def __init__(self, cameras=None, raster_settings=None):
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else RasterizationSettings()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def transform(self, meshes_world, **kwargs):
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass")
        verts_world = meshes_world.verts_padded()
        eps = kwargs.get("eps", None)
        verts_view = cameras.get_world_to_view_transform(**kwargs).transform_points(verts_world, eps=eps)
        to_ndc = cameras.get_ndc_camera_transform(**kwargs)
        proj = cameras.get_projection_transform(**kwargs).compose(to_ndc)
        verts_ndc = proj.transform_points(verts_view, eps=eps)
        verts_ndc = torch.cat([verts_ndc[..., :2], verts_view[..., 2:3]], dim=-1)
        return meshes_world.update_padded(new_verts_padded=verts_ndc)

    def forward(self, meshes_world, **kwargs):
        meshes_proj = self.transform(meshes_world, **kwargs)
        rs = kwargs.get("raster_settings", self.raster_settings)
        clip = rs.clip_barycentric_coords
        if clip is None:
            clip = rs.blur_radius > 0.0
        cameras = kwargs.get("cameras", self.cameras)
        persp = rs.perspective_correct if rs.perspective_correct is not None else cameras.is_perspective()
        S = rs.image_size if isinstance(rs.image_size, int) else rs.image_size[0]
        p2f, zbuf, bary, dists, ov = _raster.rasterize_meshes(
            meshes_proj.verts_padded(), meshes_proj.faces_padded(), S, rs.blur_radius, rs.faces_per_pixel,
            persp, clip, rs.cull_backfaces)
        MeshRasterizer.last_overflow = ov
        return Fragments(pix_to_face=p2f, zbuf=zbuf, bary_coords=bary, dists=dists)


class MeshRenderer(nn.Module):
    def __init__(self, rasterizer, shader):
        super().__init__()
        self.rasterizer, self.shader = rasterizer, shader

    def to(self, device):
        self.rasterizer.to(device)
        self.shader.to(device)
        return self

    def forward(self, meshes_world, **kwargs):
        fragments = self.rasterizer(meshes_world, **kwargs)
        return self.shader(fragments, meshes_world, **kwargs)


class ShaderBase(nn.Module):
    def __init__(self, device: Device = "cpu", cameras=None, lights=None, materials=None, blend_params=None):
        super().__init__()
        self.lights = lights if lights is not None else PointLights(device=device)
        self.materials = materials if materials is not None else Materials(device=device)
        self.cameras = cameras
        self.blend_params = blend_params if blend_params is not None else BlendParams()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        self.materials = self.materials.to(device)
        self.lights = self.lights.to(device)
        return self


class SoftSilhouetteShader(nn.Module):
    def __init__(self, blend_params=None):
        super().__init__()
        self.blend_params = blend_params if blend_params is not None else BlendParams()

    def to(self, device):
        return self

    def forward(self, fragments, meshes, **kwargs):
        colors = torch.ones_like(fragments.bary_coords)
        blend_params = kwargs.get("blend_params", self.blend_params)
        return sigmoid_alpha_blend(colors, fragments, blend_params)


# ------------------------------------------------------------------------------------------------- losses
def mesh_laplacian_smoothing(meshes, method: str = "uniform"):
    """loss/mesh_laplacian_smoothing.py, uniform weights."""
    if meshes.isempty():
        return torch.tensor([0.0], dtype=torch.float32, device=meshes.device, requires_grad=True)
    N = len(meshes)
    verts_packed = meshes.verts_packed()
    num_verts_per_mesh = meshes.num_verts_per_mesh()
    verts_packed_idx = meshes.verts_packed_to_mesh_idx()
    weights = num_verts_per_mesh.gather(0, verts_packed_idx)
    weights = 1.0 / weights.to(verts_packed.dtype)
    with torch.no_grad():
        L = meshes.laplacian_packed()
    loss = torch.sparse.mm(L, verts_packed)
    loss = loss.norm(dim=1)
    loss = loss * weights
    return loss.sum() / N


def _nc_find_pairs(edge_num):
    pairs = []
    start = 0
    for n in edge_num.tolist():
        for i in range(n):
            for j in range(i + 1, n):
                pairs.append((start + i, start + j))
        start += n
    return torch.tensor(pairs, dtype=torch.long).view(-1, 2)


def mesh_normal_consistency(meshes):
    """loss/mesh_normal_consistency.py."""
    if meshes.isempty():
        return torch.tensor([0.0], dtype=torch.float32, device=meshes.device, requires_grad=True)
    N = len(meshes)
    verts_packed = meshes.verts_packed()
    faces_packed = meshes.faces_packed()
    edges_packed = meshes.edges_packed()
    verts_packed_to_mesh_idx = meshes.verts_packed_to_mesh_idx()
    face_to_edge = meshes.faces_packed_to_edges_packed()
    E = edges_packed.shape[0]
    Fp = faces_packed.shape[0]
    with torch.no_grad():
        edge_idx = face_to_edge.reshape(Fp * 3)
        vert_idx = faces_packed.view(1, Fp, 3).expand(3, Fp, 3).transpose(0, 1).reshape(3 * Fp, 3)
        edge_idx, edge_sort_idx = edge_idx.sort()
        vert_idx = vert_idx[edge_sort_idx]
        edge_num = edge_idx.bincount(minlength=E)
        vert_edge_pair_idx = _nc_find_pairs(edge_num.cpu()).to(edge_num.device)
    if vert_edge_pair_idx.shape[0] == 0:
        return torch.tensor([0.0], dtype=torch.float32, device=meshes.device, requires_grad=True)
    v0_idx = edges_packed[edge_idx, 0]
    v0 = verts_packed[v0_idx]
    v1_idx = edges_packed[edge_idx, 1]
    v1 = verts_packed[v1_idx]
    n_temp0 = (v1 - v0).cross(verts_packed[vert_idx[:, 0]] - v0, dim=1)
    n_temp1 = (v1 - v0).cross(verts_packed[vert_idx[:, 1]] - v0, dim=1)
    n_temp2 = (v1 - v0).cross(verts_packed[vert_idx[:, 2]] - v0, dim=1)
    n = n_temp0 + n_temp1 + n_temp2
    n0 = n[vert_edge_pair_idx[:, 0]]
    n1 = -n[vert_edge_pair_idx[:, 1]]
    loss = 1 - torch.cosine_similarity(n0, n1, dim=1)
    v2m = verts_packed_to_mesh_idx[vert_idx[:, 0]]
    v2m = v2m[vert_edge_pair_idx[:, 0]]
    num_normals = v2m.bincount(minlength=N)
    weights = 1.0 / num_normals[v2m].to(verts_packed.dtype)
    loss = loss * weights
    return loss.sum() / N
