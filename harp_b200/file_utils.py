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
        params = dict(params.params())  # a HarpFitter: optimised tensors + the keys it was built with (extra_params)
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


def prepare_resumed_params(params, known_appearance_new_pose=None):
    """The resume-time re-initialisation of `optimize_sequence.py:355-390` (`start_from`), on whatever device the tensors are on:
    poses linearly re-interpolated between every 30th frame (367-371), translation and root rotation reset to their sequence
    means (373-379), and defaults for the keys older checkpoints lack (381-389).  `known_appearance_new_pose`: optional dict
    with trans / pose / rot / cam of the new sequence (359-365).  Returns a new dict of tensors (requires_grad is the
    caller's business, as in `HarpFitter`)."""
    out = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in params.items()}
    if known_appearance_new_pose is not None:
        for k in ("trans", "pose", "rot", "cam"):
            out[k] = known_appearance_new_pose[k].detach().clone()
    pose = out["pose"]
    temp_pose = pose.clone()
    for i in range(pose.shape[0] // 30 - 1):
        j = torch.arange(30, device=pose.device, dtype=pose.dtype)[:, None]
        temp_pose[i * 30:i * 30 + 30] = ((30 - j) * pose[i * 30] + j * pose[i * 30 + 30]) / 30.0
    out["pose"] = temp_pose
    out["trans"] = torch.zeros_like(out["trans"]) + out["trans"].mean(0)
    out["rot"] = torch.zeros_like(out["rot"]) + out["rot"].mean(0)
    dev = pose.device
    if out.get("wrist_pose") is None:
        out["wrist_pose"] = torch.zeros([pose.shape[0], 3], device=dev)
    if out.get("amb_ratio") is None:
        out["amb_ratio"] = torch.tensor(0.4, device=dev)  # roughly 0.6 after the sigmoid
    if out.get("normal_map") is None:
        out["normal_map"] = torch.tensor([0.0, 0.0, 1.0], device=dev).repeat(1, 512, 512, 1)
    return out
