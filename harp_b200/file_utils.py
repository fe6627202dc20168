"""Checkpoint compatibility with `utils/file_utils.py:6-37`: `saved_params[_test].pkl` is a pickle of {key: ndarray} with the
keys of `optimize_sequence.py:init_params` (181-249).  `save_result` / `load_result` read and write the same file from a
`HarpFitter` (or any dict of tensors), so a run can move between the reference and this package."""
import os
import pickle

import numpy as np
import torch

PARAM_KEYS = ("trans", "pose", "rot", "shape", "wrist_pose", "verts_disps", "verts_rgb", "texture", "normal_map",
              "light_positions", "amb_ratio", "cam", "mesh_faces", "verts_uvs", "faces_uvs", "uv_mask", "init_joints")


def save_result(params, base_output_dir, test=False):
    """utils/file_utils.py:6-16.  `params`: dict of tensors / arrays, or a HarpFitter (its `.params()` are used)."""
    if hasattr(params, "params") and callable(params.params):
        extra = getattr(params, "extra_params", {})
        params = dict(params.params(), **extra)
    out = {}
    for k, v in params.items():
        if v is None:
            out[k] = None
        elif torch.is_tensor(v):
            out[k] = v.detach().cpu().numpy()
        else:
            out[k] = np.asarray(v)
    name = "saved_params_test.pkl" if test else "saved_params.pkl"
    os.makedirs(base_output_dir, exist_ok=True)
    with open(os.path.join(base_output_dir, name), "wb") as fh:
        pickle.dump(out, fh)
    return os.path.join(base_output_dir, name)


def load_result(base_output_dir, test=False):
    """utils/file_utils.py:18-26 -> dict of torch tensors (requires_grad handled by the caller, 28-37)."""
    name = "saved_params_test.pkl" if test else "saved_params.pkl"
    with open(os.path.join(base_output_dir, name), "rb") as fh:
        raw = pickle.load(fh)
    return {k: (torch.from_numpy(np.asarray(v)) if v is not None else None) for k, v in raw.items()}
