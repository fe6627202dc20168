"""smplx/vertex_joint_selector.py (restated): append vertex-defined joints (face, feet, fingertips) to the regressed ones."""
import numpy as np
import torch
import torch.nn as nn

from .utils import to_tensor


class VertexJointSelector(nn.Module):
    def __init__(self, vertex_ids=None, use_hands=True, use_feet_keypoints=True, **kwargs):
        super().__init__()
        face = np.array([vertex_ids["nose"], vertex_ids["reye"], vertex_ids["leye"], vertex_ids["rear"], vertex_ids["lear"]], dtype=np.int64)
        extra = np.concatenate([np.array([], dtype=np.int64), face])
        if use_feet_keypoints:
            feet = np.array([vertex_ids["LBigToe"], vertex_ids["LSmallToe"], vertex_ids["LHeel"], vertex_ids["RBigToe"],
                             vertex_ids["RSmallToe"], vertex_ids["RHeel"]], dtype=np.int32)
            extra = np.concatenate([extra, feet])
        if use_hands:
            self.tip_names = ["thumb", "index", "middle", "ring", "pinky"]
            tips = [vertex_ids[h + n] for h in ["l", "r"] for n in self.tip_names]
            extra = np.concatenate([extra, tips])
        self.register_buffer("extra_joints_idxs", to_tensor(extra, dtype=torch.long))

    def forward(self, vertices, joints):
        extra_joints = torch.index_select(vertices, 1, self.extra_joints_idxs)
        return torch.cat([joints, extra_joints], dim=1)
