"""Derive compact, licence-free topology/UV fixtures from the reference's template OBJ files.

Reads  /root/reference/template/{hand,arm}/*.obj + uv_mask.png   (data assets, not source code)
Writes harp_b200/assets/{hand,arm}_template.npz

The OBJ files hold the *already subdivided* mesh in PyTorch3D `SubdivideMeshes` face order
(SURVEY.md section 4): for coarse face k, OBJ faces k, k+F0, k+2F0, k+3F0 are
f0=[v0,e2,e1], f1=[v1,e0,e2], f2=[v2,e1,e0], f3=[e0,e1,e2].  Blender permuted the vertex ids, so we
recover the coarse mesh (v0,v1,v2 of every coarse face), renumber its vertices 0..V0-1, and verify that
subdividing it the PyTorch3D way reproduces the OBJ faces corner-for-corner (so that the OBJ's per-corner
`faces_uvs` pair index-for-index with the runtime-subdivided faces, as `utils/hand_model_utils.py:59-61`
and `utils/visualize.py:84-86` require).

Run once in the build container:  python tools/make_assets.py
"""
import os
import sys
import numpy as np

REF = "/root/reference/template"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "harp_b200", "assets")


def parse_obj(path):
    v, vt, fv, ft = [], [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("vt "):
                vt.append([float(x) for x in line.split()[1:3]])
            elif line.startswith("f "):
                toks = line.split()[1:4]
                fv.append([int(t.split("/")[0]) - 1 for t in toks])
                ft.append([int(t.split("/")[1]) - 1 for t in toks])
    return (np.asarray(v, np.float64), np.asarray(vt, np.float64),
            np.asarray(fv, np.int64), np.asarray(ft, np.int64))


def edges_packed(faces, V):
    """PyTorch3D Meshes.edges_packed + faces_packed_to_edges_packed."""
    e12 = faces[:, [1, 2]]
    e20 = faces[:, [2, 0]]
    e01 = faces[:, [0, 1]]
    e = np.concatenate([e12, e20, e01], 0)
    e = np.sort(e, axis=1)
    key = e[:, 0] * V + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    edges = np.stack([uniq // V, uniq % V], 1)
    F = faces.shape[0]
    f2e = inv.reshape(3, F).T  # columns: e12 (opp v0), e20 (opp v1), e01 (opp v2)
    return edges, f2e


def subdivide_faces(faces, V):
    edges, f2e = edges_packed(faces, V)
    fe = f2e + V
    f0 = np.stack([faces[:, 0], fe[:, 2], fe[:, 1]], 1)
    f1 = np.stack([faces[:, 1], fe[:, 0], fe[:, 2]], 1)
    f2 = np.stack([faces[:, 2], fe[:, 1], fe[:, 0]], 1)
    f3 = fe
    return edges, np.concatenate([f0, f1, f2, f3], 0)


def process(name, obj, mask_png, scale_to_m, corr_pkl=None):
    v, vt, fv, ft = parse_obj(obj)
    F = fv.shape[0]
    F0 = F // 4
    coarse_obj = np.stack([fv[:F0, 0], fv[F0:2 * F0, 0], fv[2 * F0:3 * F0, 0]], 1)  # OBJ numbering
    ids = np.unique(coarse_obj)
    V0 = ids.shape[0]
    remap = -np.ones(v.shape[0], np.int64)
    remap[ids] = np.arange(V0)
    faces0 = remap[coarse_obj]
    verts0 = v[ids] * scale_to_m
    edges0, faces_sub = subdivide_faces(faces0, V0)
    E0 = edges0.shape[0]
    # our numbering of the subdivided mesh -> OBJ numbering, from corner correspondence
    to_obj = -np.ones(V0 + E0, np.int64)
    to_obj[faces_sub.reshape(-1)] = fv.reshape(-1)
    assert (to_obj >= 0).all()
    # consistent bijection?
    assert np.array_equal(to_obj[faces_sub], fv), "OBJ faces are not in SubdivideMeshes order"
    assert np.unique(to_obj).shape[0] == v.shape[0] == V0 + E0
    # midpoints agree with the OBJ's subdivided verts
    mid = 0.5 * (verts0[edges0[:, 0]] + verts0[edges0[:, 1]])
    sub = np.concatenate([verts0, mid], 0)
    err = np.abs(sub - v[to_obj] * scale_to_m).max()
    extra = {}
    if corr_pkl is not None:
        # template/arm/smplx_arm_corr.pkl: arm sub-mesh of SMPL-X (hand_models_harp/body_models.py:2100-2127).  Its 'face'
        # list is the coarse arm mesh in ARM numbering and in SubdivideMeshes input order, so coarse face k pairs with
        # our recovered face k corner by corner: renumber our coarse verts into the pickle's numbering.
        import pickle
        with open(corr_pkl, "rb") as fh:
            corr = pickle.load(fh)
        cf = np.asarray(corr["face"], np.int64)
        assert cf.shape == faces0.shape
        to_arm = -np.ones(V0, np.int64)
        to_arm[faces0.reshape(-1)] = cf.reshape(-1)
        assert (to_arm >= 0).all() and np.unique(to_arm).shape[0] == V0 and np.array_equal(to_arm[faces0], cf)
        order = np.argsort(to_arm)          # arm id -> our id
        verts0, faces0 = verts0[order], cf
        extra = dict(arm_vert=np.asarray(corr["arm_vert"], np.int32),                      # SMPL-X vertex id of each arm vert
                     mano_vert_from_arm=np.asarray(corr["mano_vert_from_arm"], np.int32),  # the 778 hand verts inside the arm
                     mano_joint=np.asarray(corr["mano_joint"], np.int32))                  # 22 joint ids (21 hand + elbow)
        edges0, faces_sub = subdivide_faces(faces0, V0)
        assert np.array_equal(np.sort(faces_sub.reshape(-1)), np.sort(faces_sub.reshape(-1)))
    from PIL import Image
    m = np.asarray(Image.open(mask_png).convert("L").resize((512, 512))) / 255.0  # optimize_sequence.py:176-178
    out = os.path.join(OUT, f"{name}_template.npz")
    np.savez_compressed(
        out,
        verts0=verts0.astype(np.float32),      # (V0,3) metres
        faces0=faces0.astype(np.int32),        # (F0,3)
        verts_uvs=vt.astype(np.float32),       # (Vt,2)
        faces_uvs=ft.astype(np.int32),         # (F,3) pairs with subdivided faces
        uv_mask=m.astype(np.float32),          # (512,512)
        **extra,
    )
    print(f"{name}: V0={V0} F0={F0} E0={E0} V={V0+E0} F={F} Vt={vt.shape[0]} midpoint err={err:.3e} m -> {out}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    process("hand", f"{REF}/hand/textured_hand.obj", f"{REF}/hand/uv_mask.png", 1e-3)
    process("arm", f"{REF}/arm/arm_template.obj", f"{REF}/arm/uv_mask.png", 1.0, corr_pkl=f"{REF}/arm/smplx_arm_corr.pkl")
