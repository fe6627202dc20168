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


# ----------------------------------------------------------------------------------------------------- SMPL-X arm
# standard SMPL-X kinematic tree (55 joints: 22 body, jaw, 2 eyes, 15 + 15 hand joints)
SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]
SMPLX_NUM_VERTS = 10475
# VertexJointSelector extra joints (smplx/vertex_ids.py, 'smplx' table): face 5, feet 6, left tips 5, right tips 5
SMPLX_VERTEX_IDS = dict(nose=9120, reye=9929, leye=9448, rear=616, lear=6, rthumb=8079, rindex=7669, rmiddle=7794, rring=7905,
                        rpinky=8022, lthumb=5361, lindex=4933, lmiddle=5058, lring=5169, lpinky=5286, LBigToe=5770,
                        LSmallToe=5780, LHeel=8846, RBigToe=8463, RSmallToe=8474, RHeel=8635)
SMPLX_RIGHT_TIPS = ["rthumb", "rindex", "rmiddle", "rring", "rpinky"]  # extra joints 71..75


def make_smplx_pack(seed=0, template=None):
    """SMPL-X-shaped model dict (field names of SMPLX_NEUTRAL.npz as read at hand_models_harp/body_models.py:143-260,
    602-623,2013-2067) that is anatomically meaningful only on the right arm: the 1026 arm vertices sit at the real
    `template/arm` positions (via `smplx_arm_corr.pkl`), everything else is filler.  float32."""
    t = template if template is not None else load_template("arm")
    rng = np.random.RandomState(seed)
    V, J = SMPLX_NUM_VERTS, 55
    arm_ids = t["arm_vert"].astype(np.int64)
    va = t["verts0"].astype(np.float64)                     # (1026,3) arm verts in arm numbering
    hand_local = t["mano_vert_from_arm"].astype(np.int64)
    v = rng.normal(0, 0.15, size=(V, 3)) + np.array([0.2, 0.0, 0.0])
    v[arm_ids] = va
    # joints: wrist / elbow from the geometry, fingers by farthest-point sampling on the hand part (as make_mano_pack)
    hand = va[hand_local]
    forearm = np.setdiff1d(np.arange(va.shape[0]), hand_local)
    wrist = hand[np.argsort(hand[:, 0])[-40:]].mean(0)     # the hand's proximal rim (largest x)
    elbow = va[forearm][np.argsort(va[forearm][:, 0])[-40:]].mean(0)
    d = np.linalg.norm(hand - wrist, axis=1)
    tips = []
    for _ in range(5):
        i = int(np.argmax(d)); tips.append(i)
        d = np.minimum(d, np.linalg.norm(hand - hand[i], axis=1) * 1.5)
    Jpos = rng.normal(0, 0.15, size=(J, 3)) + np.array([0.2, 0.0, 0.0])
    Jpos[19], Jpos[21] = elbow, wrist
    Jpos[17] = elbow + (elbow - wrist)                      # shoulder continues the forearm axis
    Jpos[14] = Jpos[17] + np.array([0.15, 0.0, 0.0])
    # right-hand joints 40..54 in SMPL-X order: index, middle, pinky, ring, thumb (3 each)
    for f in range(5):
        for k, a in enumerate((0.45, 0.65, 0.82)):
            Jpos[40 + 3 * f + k] = wrist + a * (hand[tips[f]] - wrist)
    # regressor rows: softmax over the nearest vertices (arm verts for arm joints), rows sum to 1
    Jreg = np.zeros((J, V))
    for j in range(J):
        d2 = ((v - Jpos[j]) ** 2).sum(1)
        nn_ = np.argsort(d2)[:24]
        w = np.exp(-d2[nn_] / (2 * 0.012 ** 2) + d2[nn_].min() / (2 * 0.012 ** 2)); w /= w.sum()
        Jreg[j, nn_] = w
    # skinning weights: arm verts over {elbow, wrist, right-hand joints}; everything else rigid to the pelvis
    W = np.zeros((V, J)); W[:, 0] = 1.0
    cand = np.array([19, 21] + list(range(40, 55)))
    Jtrue = Jreg @ v
    d2 = ((va[:, None, :] - Jtrue[cand][None]) ** 2).sum(-1)
    w = np.exp(-d2 / (2 * 0.025 ** 2))
    kth = np.sort(w, axis=1)[:, -4][:, None]
    w = np.where(w >= kth, w, 0.0); w /= w.sum(1, keepdims=True)
    W[arm_ids] = 0.0
    W[np.repeat(arm_ids, cand.size), np.tile(cand, arm_ids.size)] = w.reshape(-1)
    nb = _one_ring(va.shape[0], t["faces0"])

    def smooth(x, it=2):
        for _ in range(it):
            x = np.stack([(x[i] + sum(x[j] for j in nb[i])) / (1 + len(nb[i])) for i in range(x.shape[0])], 0)
        return x

    shapedirs = np.zeros((V, 3, 20), np.float32)
    shapedirs[arm_ids] = (smooth(rng.normal(0, 2e-3, size=(va.shape[0], 60))) * 3.0).reshape(-1, 3, 20)
    posedirs = np.zeros((V, 3, 486), np.float32)
    posedirs[arm_ids] = (smooth(rng.normal(0, 0.5e-3, size=(va.shape[0], 1458)), it=1) * 2.0).reshape(-1, 3, 486)
    faces = arm_ids[t["faces0"].astype(np.int64)]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(v_template=f32(v), shapedirs=shapedirs, posedirs=posedirs, J_regressor=f32(Jreg), weights=f32(W),
                kintree_table=np.stack([np.asarray(SMPLX_PARENTS, np.int64), np.arange(J)], 0), f=faces.astype(np.int64),
                hands_componentsl=np.eye(45, dtype=np.float32), hands_componentsr=np.eye(45, dtype=np.float32),
                hands_meanl=f32(rng.normal(0, 0.1, size=45)), hands_meanr=f32(rng.normal(0, 0.1, size=45)),
                lmk_faces_idx=np.zeros(51, np.int64), lmk_bary_coords=np.full((51, 3), 1.0 / 3.0, np.float32),
                dynamic_lmk_faces_idx=np.zeros((79, 17), np.int64), dynamic_lmk_bary_coords=np.full((79, 17, 3), 1.0 / 3.0, np.float32))


def reduce_smplx_arm(pack, template=None):
    """Restrict a SMPL-X model dict to what SMPLXARM returns (hand_models_harp/body_models.py:2383-2390): the 1026 arm
    vertices and the 22 `mano_joint` joints; fold the joint regressor; express fingertip vertex-joints in arm numbering.
    -> kwargs of `harp_b200.ops.HarpContext(..., lbs=...)` (the generic skinning model of include/harp_b200.h)."""
    t = template if template is not None else load_template("arm")
    arm_ids = t["arm_vert"].astype(np.int64)
    inv = -np.ones(SMPLX_NUM_VERTS, np.int64); inv[arm_ids] = np.arange(arm_ids.size)
    n_shape = 10  # betas; the 10 expression coefficients are the layer's zero default (body_models.py:2248, 2323)
    Jreg = pack["J_regressor"].astype(np.float64)
    J_template = Jreg @ pack["v_template"].astype(np.float64)
    J_shapedirs = np.einsum("jv,vkb->jkb", Jreg, pack["shapedirs"][:, :, :n_shape].astype(np.float64))
    tips_smplx = [SMPLX_VERTEX_IDS[n] for n in SMPLX_RIGHT_TIPS]
    tips = inv[tips_smplx]
    assert (tips >= 0).all(), "right-hand fingertip vertices must belong to the arm sub-mesh"
    pose_mean = np.concatenate([np.zeros(3 * 25), pack["hands_meanl"], pack["hands_meanr"]])  # create_mean_pose, 2129-2147
    src = []
    for j in t["mano_joint"].tolist():  # 55 chain joints, then VertexJointSelector extras: right tips are 71..75
        src.append(j if j < 55 else 55 + (j - 71))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(J=55, n_shape=n_shape, parents=np.asarray(SMPLX_PARENTS, np.int32), shapedirs=f32(pack["shapedirs"][arm_ids][:, :, :n_shape]),
                posedirs=f32(pack["posedirs"][arm_ids]), v_template=f32(pack["v_template"][arm_ids]), J_template=f32(J_template),
                J_shapedirs=f32(J_shapedirs), weights=f32(pack["weights"][arm_ids]), pose_mean=f32(pose_mean), rodrigues_kind=1,
                center_joint=21, joint_out_src=np.asarray(src, np.int32), tips=tips.astype(np.int32), faces0=t["faces0"].astype(np.int32))
