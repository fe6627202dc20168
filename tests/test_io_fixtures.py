"""File-facing rows of SURVEY.md section 8 against fixtures the REFERENCE'S OWN functions produced (tests/golden/make_golden_io.py):
f1 `utils/data_util.py:load_img` (PNG -> float frames, cv2.erode'd mask), f3 `utils/file_utils.py` checkpoints."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
IO = os.path.join(HERE, "golden", "io")


def _pngs(i):
    rgb = np.asarray(Image.open(os.path.join(IO, f"frame_{i}.png")).convert("RGB"))
    mask = np.asarray(Image.open(os.path.join(IO, f"mask_{i}.png")).convert("L"))
    return rgb, mask


def test_oracle_erosion_is_cv2_erode_of_the_reference():
    """The CPU restatement of load_img's erosion (synth.erode_mask, used to build every synthetic target) against the
    reference's own load_img output on real PNG files, and against cv2.erode directly when cv2 is importable."""
    from harp_b200 import synth
    ref = np.load(os.path.join(IO, "load_img.npz"))
    for i in range(2):
        rgb, mask = _pngs(i)
        assert np.array_equal(torch.Tensor(rgb / 255).numpy(), ref[f"col_{i}"])
        m = torch.Tensor(mask[:, :, None] / 255)
        assert np.array_equal(m.numpy(), ref[f"mask_{i}"])
        er = synth.erode_mask(m[None, :, :, 0])[0].numpy()
        assert np.array_equal(er, ref[f"eroded_{i}"]), "erode_mask differs from the reference's cv2.erode(3x3, iterations=2)"
        try:
            import cv2
        except ImportError:
            continue
        assert np.array_equal(cv2.erode(mask / 255, np.ones((3, 3), np.uint8), iterations=2).astype(np.float32), ref[f"eroded_{i}"])


@pytest.mark.gpu
def test_ingest_u8_matches_the_references_load_img():
    """harp_ingest_u8 (uint8 frames + masks -> float frames, eroded mask, on the device) is bit-identical to what the
    reference's DataLoader yields for the same PNG files."""
    from harp_b200.ops import ingest_u8
    ref = np.load(os.path.join(IO, "load_img.npz"))
    rgb = np.stack([_pngs(i)[0] for i in range(2)])
    mask = np.stack([_pngs(i)[1] for i in range(2)])
    out = ingest_u8(torch.tensor(rgb).cuda(), torch.tensor(mask).cuda())
    for i in range(2):
        assert np.array_equal(out["y_true"][i].cpu().numpy(), ref[f"col_{i}"])
        assert np.array_equal(out["y_sil"][i].cpu().numpy(), ref[f"mask_{i}"][:, :, 0])
        assert np.array_equal(out["y_sil_eroded"][i].cpu().numpy(), ref[f"eroded_{i}"])


def test_checkpoint_written_by_the_reference_is_read():
    from harp_b200 import file_utils
    with open(os.path.join(IO, "ref", "saved_params.pkl"), "rb") as fh:
        raw = pickle.load(fh)
    got = file_utils.load_result(os.path.join(IO, "ref"))
    assert set(got) == set(raw) >= {"trans", "pose", "rot", "shape", "cam", "verts_disps", "texture", "normal_map", "light_positions",
                                    "amb_ratio", "mesh_faces", "verts_uvs", "faces_uvs", "uv_mask", "init_joints", "wrist_pose"}
    for k, v in raw.items():
        assert np.array_equal(got[k].numpy(), v), k


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container)")
def test_checkpoint_round_trip_through_the_references_own_functions(tmp_path):
    """Live, both directions: the reference's load_result reads what this package writes (from a dict with every key
    init_params creates), and this package reads what the reference's save_result writes."""
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from utils import file_utils as ref_fu
    from harp_b200 import file_utils as our_fu
    from make_golden_io import params_dict
    P = params_dict(seed=3)
    our_fu.save_result(P, str(tmp_path / "ours"))
    ref_fu.set_require_grad.__defaults__ = ("cpu",)  # its default device is 'cuda' (file_utils.py:28); the body runs unmodified
    back = ref_fu.load_result(str(tmp_path / "ours"), device="cpu")
    assert set(back) == set(P)
    for k in P:
        assert torch.equal(back[k].detach(), P[k]), k
    os.makedirs(tmp_path / "ref")
    ref_fu.save_result(P, str(tmp_path / "ref"))
    mine = our_fu.load_result(str(tmp_path / "ref"))
    assert all(torch.equal(mine[k], P[k]) for k in P)


def test_per_frame_fit_files_written_by_the_reference_are_read(tmp_path):
    """`%04d_mano.pkl` (hand_utils.py:766-782 writer; hand_utils.py:384-398 / data_util.py:129-147 readers): the files in
    tests/golden/io/mano_ref were written by the reference's own write_pkl; this package's writer must produce the same bytes'
    worth of content and its readers the same tensors."""
    from harp_b200 import data_util
    d_ref = os.path.join(IO, "mano_ref")
    got = data_util.load_params(d_ref)
    assert got["cam"].shape == (4, 3) and got["pose"].shape == (4, 45) and got["joints"].shape == (4, 21, 3)
    data_util.write_pkl(got, str(tmp_path / "seq" / "metro_mano"))
    for name in sorted(os.listdir(d_ref)):
        with open(os.path.join(d_ref, name), "rb") as a, open(tmp_path / "seq" / "metro_mano" / name, "rb") as b:
            ra, rb = pickle.load(a), pickle.load(b)
        assert set(ra) == set(rb)
        for k in ra:
            assert ra[k].shape == rb[k].shape and ra[k].dtype == rb[k].dtype and np.array_equal(ra[k], rb[k]), (name, k)
    merged, names = data_util.load_mano_sequences(str(tmp_path), ["seq"], average_cam_sequence=True)
    assert [n for _, n in names] == ["0001", "0002", "0003", "0004"] and merged["seq"] == ["seq"] * 4
    assert torch.allclose(merged["cam"], got["cam"].mean(0).expand(4, 3))
