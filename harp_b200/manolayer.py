"""Drop-in for `manopth.manolayer.ManoLayer` (use_pca=False, axis-angle, right hand) backed by `lbs_mano_{fwd,bwd}`.

The licensed `MANO_RIGHT.pkl` cannot ship, so the layer is built from a dict holding the buffers the reference registers
(`manopth/manolayer.py:70-97`: th_shapedirs, th_posedirs, th_v_template, th_J_regressor, th_weights, th_hands_mean,
th_faces) plus the five fingertip vertex ids (`manolayer.py:272`)."""
import numpy as np
import torch

from . import ops, structures


class ManoLayer(torch.nn.Module):
    def __init__(self, model, faces_uvs=None, verts_uvs=None, device="cuda:0"):
        super().__init__()
        for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights", "th_hands_mean"):
            if k in model:
                self.register_buffer(k, torch.as_tensor(np.asarray(model[k]), dtype=torch.float32))
        self.register_buffer("th_faces", torch.as_tensor(np.asarray(model["th_faces"])).long())
        self.side, self.use_pca, self.ncomps, self.rot, self.center_idx = "right", False, 45, 3, None
        # one context per topology the loop uses: coarse (mesh_subdivider=None) and SubdivideMeshes'd
        self.hc_coarse = structures.register_context(ops.HarpContext(model["th_faces"], subdivide=False, faces_uvs=faces_uvs,
                                                                     verts_uvs=verts_uvs, mano=model, device=device))
        self.hc = structures.register_context(ops.HarpContext(model["th_faces"], subdivide=True, faces_uvs=faces_uvs,
                                                              verts_uvs=verts_uvs, mano=model, device=device))
        self.to(device)

    def forward(self, th_pose_coeffs, th_betas=torch.zeros(1), th_trans=torch.zeros(1), root_palm=torch.Tensor([0]),
                share_betas=torch.Tensor([0]), no_root_rot=False):
        """(B,48) pose coefficients [rot | pose], (B,10) betas, (B,3) trans -> (verts mm (B,778,3), joints mm (B,21,3))."""
        if bool(root_palm) or bool(share_betas) or no_root_rot:
            raise NotImplementedError("root_palm / share_betas / no_root_rot are not used by HARP (utils/visualize.py:42-44)")
        B = th_pose_coeffs.shape[0]
        dev = th_pose_coeffs.device
        if th_betas is None or th_betas.numel() == 1:
            th_betas = self.th_betas.expand(B, 10) if hasattr(self, "th_betas") else torch.zeros(B, 10, device=dev)
        if th_trans is None or th_trans.numel() == 1:
            th_trans = torch.zeros(B, 3, device=dev)
        verts, joints = ops.LbsMano.apply(self.hc, th_pose_coeffs, th_betas, th_trans)
        return verts * 1000, joints * 1000  # manolayer.py:291-292
