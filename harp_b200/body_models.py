"""Drop-in for `hand_models_harp.body_models.SMPLXARM` as HARP calls it (`utils/visualize.py:37-40`):

    hand_verts_mm, hand_joints_mm = layer(betas=, global_orient=, transl=, right_hand_pose=, right_wrist_pose=,
                                          return_type='mano_w_arm')

Only the 1026 arm vertices and 22 joints the reference returns are skinned (the reference skins all 10 475 SMPL-X vertices
and drops 90 %: body_models.py:2335-2339, 2383).  Built from a SMPL-X-format model dict (the fields of SMPLX_NEUTRAL.npz)
plus the arm correspondence (`template/arm/smplx_arm_corr.pkl`, shipped as harp_b200/assets/arm_template.npz)."""
import torch

from . import ops, structures, synth


class SMPLXARM(torch.nn.Module):
    NUM_BODY_JOINTS, NUM_HAND_JOINTS = 21, 15

    def __init__(self, smplx_model, arm_template=None, device="cuda:0"):
        super().__init__()
        t = arm_template if arm_template is not None else synth.load_template("arm")
        red = synth.reduce_smplx_arm(smplx_model, t)
        self.right_arm_faces_tensor = torch.as_tensor(t["faces0"]).long()  # body_models.py:2113-2116
        self.right_mano_idx = torch.as_tensor(t["mano_vert_from_arm"]).long().to(device)  # MANO's 778 vertices inside the arm mesh
        kw = dict(faces_uvs=t["faces_uvs"], verts_uvs=t["verts_uvs"], lbs=red, device=device)
        self.hc_coarse = structures.register_context(ops.HarpContext(t["faces0"], subdivide=False, **kw))
        self.hc = structures.register_context(ops.HarpContext(t["faces0"], subdivide=True, **kw))
        self.device = torch.device(device)

    @staticmethod
    def assemble_full_pose(global_orient, right_hand_pose, right_wrist_pose=None):
        """(B,165) axis-angle pose in SMPL-X joint order with the layer's zero defaults everywhere else
        (body_models.py:2299-2311): [global(3) | body 21x3 with joint 21 = right wrist | jaw, eyes 9 | left hand 45 | right hand 45]."""
        B = global_orient.shape[0]
        z = lambda n: global_orient.new_zeros(B, n)
        wrist = right_wrist_pose if right_wrist_pose is not None else z(3)
        return torch.cat([global_orient, z(60), wrist, z(9), z(45), right_hand_pose], dim=1)

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None, right_hand_pose=None, transl=None,
                expression=None, jaw_pose=None, leye_pose=None, reye_pose=None, return_type="mano_w_arm", hc=None, **kwargs):
        for name, v in (("body_pose", body_pose), ("left_hand_pose", left_hand_pose), ("expression", expression),
                        ("jaw_pose", jaw_pose), ("leye_pose", leye_pose), ("reye_pose", reye_pose)):
            if v is not None:
                raise NotImplementedError(f"SMPLXARM drop-in: `{name}` stays at the layer default in HARP (utils/visualize.py:37-40)")
        if return_type not in ("mano_w_arm", "mano"):
            raise NotImplementedError("return_type must be 'mano_w_arm' (fitting loop) or 'mano' (upstream smoothing)")
        B = global_orient.shape[0]
        transl = transl if transl is not None else global_orient.new_zeros(B, 3)
        right_hand_pose = right_hand_pose if right_hand_pose is not None else global_orient.new_zeros(B, 45)
        full_pose = self.assemble_full_pose(global_orient, right_hand_pose, kwargs.get("right_wrist_pose"))
        verts, joints = ops.Lbs.apply(hc or self.hc, full_pose, betas, transl)
        if return_type == "mano":  # body_models.py:2387-2388: the hand's 778 vertices in MANO order, the 21 MANO joints
            return verts[:, self.right_mano_idx] * 1000.0, joints[:, :21] * 1000.0
        return verts * 1000.0, joints * 1000.0  # body_models.py:2383-2385
