"""Synthetic, seeded stand-ins for the licensed assets and the per-sequence inputs of the HARP hot path.

The reference needs `mano/models/MANO_RIGHT.pkl` (`manopth/manolayer.py:60-65`) and per-frame METRO outputs
(`utils/data_util.py:129-147`); neither can ship.  This module builds MANO-*shaped* buffers on the real template
topology (`harp_b200/assets/hand_template.npz`, derived by `tools/make_assets.py` from the reference's
`template/hand/textured_hand.obj`) following SURVEY.md section 8(d), plus per-frame parameters and appearance maps.
Everything is numpy + torch on CPU and deterministic in the seed; the CUDA path and the oracle both consume it.
"""
import os

import numpy as np
import torch

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

# manopth kinematic tree as used by ManoLayer.forward (manolayer.py:209-242): joint j's parent
MANO_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
MANO_TIP_IDS = [745, 317, 444, 556, 673]  # manolayer.py:272 (right hand)
TEX_SIZE = 512  # optimize_sequence.py:233 (texture and normal map are (1,512,512,3))


def load_template(name="hand"):
    d = np.load(os.path.join(ASSET_DIR, f"{name}_template.npz"))
    return {k: d[k] for k in d.files}


def _one_ring(V, faces):
    nb = [set() for _ in range(V)]
    for a, b, c in faces:
        nb[a].update((b, c)); nb[b].update((a, c)); nb[c].update((a, b))
    return nb


def make_mano_pack(seed=0, template=None):
    """MANO-shaped buffers (float32 numpy), names as registered at manolayer.py:70-97."""
    t = template if template is not None else load_template("hand")
    rng = np.random.RandomState(seed)
    v = t["verts0"].astype(np.float64)  # (778,3) metres
    faces = t["faces0"].astype(np.int64)
    V = v.shape[0]
    nb = _one_ring(V, faces)

    def smooth(x, it=2):
        for _ in range(it):
            x = np.stack([(x[i] + sum(x[j] for j in nb[i])) / (1 + len(nb[i])) for i in range(V)], 0)
        return x

    shapedirs = smooth(rng.normal(0, 2e-3, size=(V, 30))).reshape(V, 3, 10) * 3.0
    posedirs = smooth(rng.normal(0, 0.5e-3, size=(V, 405)), it=1).reshape(V, 3, 135) * 2.0
    # wrist = centroid of the open boundary loop; finger tips by farthest-point sampling
    ecount = {}
    for f in faces:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            k = (min(a, b), max(a, b)); ecount[k] = ecount.get(k, 0) + 1
    bverts = sorted({i for k, c in ecount.items() if c == 1 for i in k})
    wrist = v[bverts].mean(0) if bverts else v.mean(0)
    d = np.linalg.norm(v - wrist, axis=1)
    tips = []
    for _ in range(5):
        i = int(np.argmax(d)); tips.append(i)
        d = np.minimum(d, np.linalg.norm(v - v[i], axis=1) * 1.5)
    J = [wrist]
    for tip in tips:
        for a in (0.45, 0.65, 0.82):
            J.append(wrist + a * (v[tip] - wrist))
    J = np.stack(J, 0)  # (16,3)
    d2 = ((J[:, None, :] - v[None, :, :]) ** 2).sum(-1)  # (16,V)
    Jreg = np.exp(-d2 / (2 * 0.012 ** 2)); Jreg /= Jreg.sum(1, keepdims=True)
    w = np.exp(-d2.T / (2 * 0.02 ** 2))  # (V,16)
    kth = np.sort(w, axis=1)[:, -4][:, None]
    w = np.where(w >= kth, w, 0.0); w /= w.sum(1, keepdims=True)
    hands_mean = rng.normal(0, 0.1, size=(45,))
    # Renumber vertices so that the reference's hard-coded MANO fingertip ids (manolayer.py:272) select our tips.
    perm = np.arange(V)  # new id -> old id
    for h, t_ in zip(MANO_TIP_IDS, tips):
        i = int(np.where(perm == t_)[0][0])
        perm[[i, h]] = perm[[h, i]]
    inv = np.empty_like(perm); inv[perm] = np.arange(V)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(th_shapedirs=f32(shapedirs[perm]), th_posedirs=f32(posedirs[perm]), th_v_template=f32(v[perm])[None],
                th_J_regressor=f32(Jreg[:, perm]), th_weights=f32(w[perm]), th_hands_mean=f32(hands_mean)[None],
                th_betas=np.zeros((1, 10), np.float32), th_faces=inv[faces].astype(np.int64),
                tips=np.asarray(MANO_TIP_IDS, np.int64))


def make_sequence(T, seed=1, img_size=512):
    """Per-frame parameters: smooth pose, 0.2-rad root wobble, weak-perspective cam."""
    rng = np.random.RandomState(seed)
    t = np.arange(T)[:, None] / max(T, 1)
    k = rng.randint(1, 4, size=(1, 45)); ph = rng.uniform(0, 2 * np.pi, size=(1, 45))
    pose = 0.25 * np.sin(2 * np.pi * k * t + ph)
    kr = rng.randint(1, 3, size=(1, 3)); pr = rng.uniform(0, 2 * np.pi, size=(1, 3))
    rot = 0.2 * np.sin(2 * np.pi * kr * t + pr) + np.array([[0.3, -0.2, 0.1]])
    trans = np.zeros((T, 3))
    shape = rng.normal(0, 0.5, size=(10,))
    cam = np.concatenate([10.0 + 0.5 * np.sin(2 * np.pi * t + 0.3),
                          0.02 * np.sin(2 * np.pi * t + 1.0), 0.02 * np.cos(2 * np.pi * 2 * t)], 1)
    f32 = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32))
    return dict(pose=f32(pose), rot=f32(rot), trans=f32(trans), shape=f32(shape), cam=f32(cam))


def make_appearance(seed=2, size=TEX_SIZE):
    """Albedo = skin colour (optimize_sequence.py:233) + band-limited noise; normal map = +z + noise; light, amb."""
    g = torch.Generator().manual_seed(seed)

    def smooth_noise(c, amp):
        lo = torch.randn(1, c, size // 16, size // 16, generator=g)
        return amp * torch.nn.functional.interpolate(lo, size=(size, size), mode="bicubic", align_corners=False)[0].permute(1, 2, 0)

    tex = (torch.tensor([232.0, 190.0, 172.0]) / 255.0).view(1, 1, 3) + smooth_noise(3, 0.08)
    nm = torch.tensor([0.0, 0.0, 1.0]).view(1, 1, 3) + smooth_noise(3, 0.05)
    return dict(texture=tex.clamp(0.02, 0.98)[None].contiguous(), normal_map=nm[None].contiguous(),
                light_position=torch.tensor([-0.5, -0.5, -0.5]), amb_ratio=torch.tensor(0.4))


def erode_mask(mask):
    """3x3 erosion, two iterations (utils/data_util.py:18-20) on a (B,S,S) {0,1} float mask; border value = 1e30 (cv2 default)."""
    m = mask[:, None]
    for _ in range(2):
        m = -torch.nn.functional.max_pool2d(-torch.nn.functional.pad(m, (1, 1, 1, 1), value=1.0), 3, 1)
    return m[:, 0]


def make_reg_offsets(seed=3, std=1.0, size=TEX_SIZE):
    """Random neighbour offsets of `albedo_reg` / `smooth_texture_reg` (loss/texture_reg.py:15,51): the reference draws
    them from the CPU RNG each call; for parity they are an explicit input. Returns int32 (size,size,2)."""
    g = torch.Generator().manual_seed(seed)
    return torch.normal(mean=0.0, std=std, size=(size, size, 2), generator=g).to(torch.int32)
