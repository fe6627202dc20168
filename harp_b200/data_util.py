"""The per-frame fit files either side of the hot path: `%04d_mano.pkl`, written by the upstream stage
(`metro_modifications/hand_utils.py:write_pkl`, 766-782) and read back by the sequence smoothing (`load_params`, 384-398) and by
the optimiser's data loading (`utils/data_util.py:129-147, 176-190`; `combine_dict_to_batch`, 52-70).  Same file layout, same
return values: a run can hand its intermediate files between the reference and this package.
Host-side glue: nothing here touches the GPU."""
import os
import pickle

import numpy as np
import torch


def write_pkl(params, pkl_out_dir, unscreen=True):
    """hand_utils.py:766-782: one pickle per frame, {key: array[None]} except 'cam' (stored without the leading axis).
    File numbers start at 1 when the frames came from `unscreen` (the reference's default)."""
    os.makedirs(pkl_out_dir, exist_ok=True)
    n = len(params["cam"])
    for idx in range(n):
        out = {}
        for k, v in params.items():
            a = v[idx].detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v[idx])
            out[k] = a if k == "cam" else a[None]
        with open(os.path.join(pkl_out_dir, "%04d_mano.pkl" % (idx + (1 if unscreen else 0))), "wb") as fh:
            pickle.dump(out, fh)


def combine_dict_to_batch(mano_dicts):
    """utils/data_util.py:52-70 (== hand_utils.py:merge_dict, 363-381): list of per-frame dicts -> dict of stacked tensors;
    'seq' stays a list of names."""
    keys = mano_dicts[0].keys()
    out = {k: [] for k in keys}
    for d in mano_dicts:
        for k in keys:
            if k == "seq":
                out[k].append(d[k])
            elif k == "cam":
                out[k].append(torch.from_numpy(np.asarray(d[k])))
            else:
                out[k].append(torch.from_numpy(np.asarray(d[k]).squeeze(0)))
    return {k: (v if k == "seq" else torch.stack(v)) for k, v in out.items()}


merge_dict = combine_dict_to_batch


def load_params(mano_dir):
    """hand_utils.py:384-398: every `*.pkl` of a directory, in name order, merged."""
    names = sorted(os.path.join(mano_dir, n) for n in os.listdir(mano_dir) if n.endswith(".pkl"))
    frames = []
    for name in names:
        with open(name, "rb") as fh:
            frames.append(pickle.load(fh))
    return combine_dict_to_batch(frames)


def load_mano_sequences(metro_output_dir, seqs, use_smooth_seq=False, average_cam_sequence=False, model_type="harp"):
    """The per-frame-fit half of `load_multiple_sequences` (utils/data_util.py:73-190): the `%04d_mano.pkl` files of the listed
    sequence folders in sorted (sequence, frame) order, each tagged with its sequence, optionally with the camera forced to the
    sequence mean (176-190).  Returns (merged dict of tensors + 'seq' names, list of (sequence, frame name))."""
    folder = ("nimble_" if model_type == "nimble" else "") + ("metro_mano_smooth" if use_smooth_seq else "metro_mano")
    names = sorted((seq, f[:-9]) for seq in seqs for f in os.listdir(os.path.join(metro_output_dir, seq, folder)) if f.endswith(".pkl"))
    frames, cams = [], {}
    for seq, name in names:
        with open(os.path.join(metro_output_dir, seq, folder, name + "_mano.pkl"), "rb") as fh:
            d = pickle.load(fh)
        d["seq"] = seq
        cams.setdefault(seq, []).append(d["cam"])
        frames.append(d)
    if average_cam_sequence:
        mean = {s: np.mean(c, axis=0) for s, c in cams.items()}
        for d in frames:
            d["cam"] = mean[d["seq"]]
    return combine_dict_to_batch(frames), names


def remove_spike(params):
    """hand_utils.py:785-801: a frame whose pose jumps by more than 1.0 (L2) from BOTH neighbours is replaced by their mean."""
    poses = params["pose"]
    step = torch.norm(poses[1:] - poses[:-1], dim=1)
    new = poses.clone().detach()
    for idx in range(1, len(poses) - 1):
        if step[idx - 1] > 1.0 and step[idx] > 1.0:
            new[idx] = (poses[idx - 1] + poses[idx + 1]) / 2.0
    params["pose"] = new
    return params
