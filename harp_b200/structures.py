"""Minimal stand-ins for the PyTorch3D containers that cross the reference's call boundary.

Only what `optimize_sequence.py:446-579` passes around: `Meshes(verts, faces, textures)` and `TexturesUV(maps, faces_uvs,
verts_uvs)`.  They hold tensors; all arithmetic happens in the CUDA kernels reached through `harp_b200.ops`.
"""
import torch

from . import ops

_CONTEXTS = {}
_VERIFIED = set()


def register_context(hc):
    """Make a `HarpContext` discoverable by mesh size, so `Meshes(verts, faces)` built by user code finds its tables."""
    _CONTEXTS.setdefault((hc.V, hc.F, str(hc.device)), []).append(hc)
    return hc


def context_for(verts, faces, textures=None):
    """The registered context whose face table EQUALS `faces` (same counts are not enough: another mesh with as many vertices
    and faces must not get this one's edges and UV tables).  The comparison runs once per (context, faces tensor)."""
    V, F = verts.shape[-2], faces.shape[-2]
    f = faces[0] if faces.dim() == 3 else faces
    cands = [hc for (v_, f_, _), lst in _CONTEXTS.items() if (v_, f_) == (V, F) for hc in lst
             if hc.device.type == verts.device.type and (hc.device.index or 0) == (verts.device.index or 0)]
    for hc in cands:
        tag = (id(hc), f.data_ptr(), f._version)
        if tag in _VERIFIED:
            return hc
        if torch.equal(hc.faces(), f.to(hc.faces().device).long()):
            _VERIFIED.add(tag)
            return hc
    # unknown topology: build tables from the faces as given (no subdivision, no skinning model)
    fu = textures.faces_uvs_first().cpu().numpy() if textures is not None else None
    vu = textures.verts_uvs_first().cpu().numpy() if textures is not None else None
    return register_context(ops.HarpContext(f.cpu().numpy(), subdivide=False, faces_uvs=fu, verts_uvs=vu, device=verts.device))


def any_context(device, n_joints=None):
    """A context on `device` for work that needs no topology (the keypoint term): optionally one whose skinning model emits
    `n_joints` joints."""
    want = torch.device(device)
    same = lambda a, b: a.type == b.type and (a.index or 0) == (b.index or 0)  # 'cuda' is 'cuda:0' unless a device was selected
    for lst in _CONTEXTS.values():
        for hc in lst:
            if same(hc.device, want) and (n_joints is None or max(hc.n_joints_out, 21) == n_joints):
                return hc
    return None


class TexturesUV:
    """`pytorch3d.renderer.TexturesUV` as used at `utils/visualize.py:84-86,99-101` (bilinear, align_corners, border)."""

    def __init__(self, maps, faces_uvs, verts_uvs):
        self._maps, self._faces_uvs, self._verts_uvs = maps, faces_uvs, verts_uvs

    def to(self, device):
        return TexturesUV(self._maps.to(device), self._faces_uvs.to(device), self._verts_uvs.to(device))

    def maps_padded(self):
        return self._maps

    def faces_uvs_first(self):
        return self._faces_uvs[0] if self._faces_uvs.dim() == 3 else self._faces_uvs

    def verts_uvs_first(self):
        return self._verts_uvs[0] if self._verts_uvs.dim() == 3 else self._verts_uvs


class TexturesVertex:
    """Vertex-colour textures (`verts_textures=True`, hard-wired off at `optimize_sequence.py:328`): container only."""

    def __init__(self, verts_features):
        self._feats = verts_features

    def to(self, device):
        return TexturesVertex(self._feats.to(device))


class Meshes:
    """Homogeneous batch of meshes: verts (B,V,3) [differentiable], faces (B,F,3) or (F,3)."""

    def __init__(self, verts, faces, textures=None):
        if isinstance(verts, (list, tuple)):
            verts = torch.stack(list(verts), 0)
        if isinstance(faces, (list, tuple)):
            faces = torch.stack(list(faces), 0)
        if verts.dim() == 2:
            verts = verts[None]
        if faces.dim() == 2:
            faces = faces[None].expand(verts.shape[0], -1, -1)
        self._verts, self._faces, self.textures = verts, faces, textures
        self.device = verts.device
        self._hc = None

    @property
    def hc(self):
        if self._hc is None:
            self._hc = context_for(self._verts, self._faces, self.textures if isinstance(self.textures, TexturesUV) else None)
        return self._hc

    def __len__(self):
        return self._verts.shape[0]

    def __getitem__(self, i):
        if isinstance(i, int):
            i = slice(i, i + 1)
        return Meshes(self._verts[i], self._faces[i], self.textures)

    def isempty(self):
        return self._verts.shape[0] == 0 or self._verts.shape[1] == 0

    def verts_padded(self):
        return self._verts

    def faces_padded(self):
        return self._faces

    def verts_packed(self):
        return self._verts.reshape(-1, 3)

    def num_verts_per_mesh(self):
        return torch.full((len(self),), self._verts.shape[1], dtype=torch.long, device=self.device)

    def num_faces_per_mesh(self):
        return torch.full((len(self),), self._faces.shape[1], dtype=torch.long, device=self.device)

    def verts_normals_padded(self):
        return ops.VertexNormals.apply(self.hc, self._verts)

    def verts_normals_packed(self):
        return self.verts_normals_padded().reshape(-1, 3)

    def extend(self, N):
        return Meshes(self._verts.repeat_interleave(N, 0), self._faces.repeat_interleave(N, 0), self.textures)
