"""Upstream stage: the per-frame MANO fit to METRO's predicted vertices,
`metro_modifications/hand_utils.py:16-133` (`optimize_for_mano_param`), on the B200 path.

The reference runs, per batch of frames, 500 Adam steps on [rot, trans] (lr 1e-1) and 700 on [rot, pose, shape, trans]
(lr 1e-2) of `MSELoss(ManoLayer(...)[0], 1000 * pred_vertices)`, re-starting (up to 4 tries) while the final loss exceeds 10.
Each step there is ~40 ATen launches forward, as many backward, and a CPU-driven optimiser.  Here one step is four
launches — `harp_lbs_mano_fwd`, `harp_mse`, `harp_lbs_mano_bwd`, `harp_adam_dev` (+ the gradient mask of the coarse stage) —
captured once per stage as a CUDA graph and replayed; the host reads the loss once per try.

Parameters live in ONE flat fp32 buffer `[pose48 (B,48) = rot|pose | trans (B,3) | shape (B,10)]` with matching grad and
Adam-moment buffers, so a stage is a single fused Adam over everything; the coarse stage zeroes the pose / shape gradients,
for which Adam with zero moments yields exactly zero update — bit-identical to an optimiser over [rot, trans] only.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream


class _FitState:
    def __init__(self, hc, target_mm):
        dev = target_mm.device
        B = target_mm.shape[0]
        self.hc, self.B, self.dev = hc, B, dev
        self.target = target_mm.contiguous()
        n = B * (48 + 3 + 10)
        self.data, self.grad = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self.m, self.v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self.state = torch.zeros(3, dtype=torch.int32, device=dev)
        o1, o2 = B * 48, B * 51
        self.pose48, self.trans, self.shape = self.data[:o1].view(B, 48), self.data[o1:o2].view(B, 3), self.data[o2:].view(B, 10)
        self.d_pose48, self.d_trans, self.d_shape = self.grad[:o1].view(B, 48), self.grad[o1:o2].view(B, 3), self.grad[o2:].view(B, 10)
        mask = torch.zeros(n, device=dev)
        mask[:o1].view(B, 48)[:, :3] = 1.0  # rot
        mask[o1:o2] = 1.0                   # trans
        self.coarse_mask = mask
        self.verts0 = torch.empty(B, hc.V0, 3, device=dev)
        self.joints = torch.empty(B, 21, 3, device=dev)
        self.d_verts0 = torch.empty(B, hc.V0, 3, device=dev)
        self.loss = torch.zeros(1, device=dev)

    def reset(self, trans_init):
        for t in (self.data, self.m, self.v):
            t.zero_()
        self.state.zero_()
        self.trans.copy_(trans_init)

    def new_optimizer(self):
        self.m.zero_()
        self.v.zero_()
        self.state.zero_()

    def forward(self):
        check(_lib.lib().harp_lbs_mano_fwd(self.hc.h, self.B, ptr(self.pose48), ptr(self.shape), 0, ptr(self.trans), ptr(self.verts0),
                                           ptr(self.joints), stream()))

    def iteration(self, lr, coarse):
        L, st = _lib.lib(), stream()
        self.loss.zero_()
        self.forward()
        check(L.harp_mse(self.verts0.numel(), ptr(self.verts0), ptr(self.target), 1000.0, ptr(self.loss), ptr(self.d_verts0), st))
        check(L.harp_lbs_mano_bwd(self.hc.h, self.B, ptr(self.pose48), ptr(self.shape), 0, ptr(self.trans), ptr(self.d_verts0), None,
                                  ptr(self.d_pose48), ptr(self.d_shape), ptr(self.d_trans), st))
        if coarse:
            self.grad.mul_(self.coarse_mask)
        check(L.harp_adam_dev(ptr(self.data), ptr(self.grad), ptr(self.m), ptr(self.v), self.data.numel(), float(lr), 0.9, 0.999, 1e-8,
                              self.state.data_ptr(), st))

    def capture(self, lr, coarse):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # one eager pass to warm up; its effect is undone by the caller's reset
            self.iteration(lr, coarse)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.iteration(lr, coarse)
        return g


def optimize_for_mano_param(pred_vertices, mano_layer, idx=0, vert_rgb=None, epoch_coarse=500, epoch_fine=700, max_tries=4,
                            use_graph=True):
    """Same call and return value as `metro_modifications/hand_utils.py:16` (epochs exposed as keyword arguments with the
    reference's values as defaults).  pred_vertices: (B,778,3) metres; mano_layer: `harp_b200.manolayer.ManoLayer`.
    Returns numpy arrays: joints (B,21,3) mm, verts (B,778,3) mm, rot (B,3), pose (B,45), shape (B,10), trans (B,3)."""
    hc = mano_layer.hc_coarse
    dev = hc.device
    pred = torch.as_tensor(pred_vertices, dtype=torch.float32).to(dev)
    if pred.dim() != 3 or pred.shape[1:] != (hc.V0, 3):
        raise ValueError(f"pred_vertices must be (B,{hc.V0},3), got {tuple(pred.shape)}")
    target = pred.detach() * 1000.0                      # hand_utils.py:39
    trans_init = target.mean(1) / 1000.0                 # 40, 62
    fs = _FitState(hc, target)
    graphs = {}
    if use_graph:
        fs.reset(trans_init)
        graphs["coarse"] = fs.capture(1e-1, True)
        graphs["fine"] = fs.capture(1e-2, False)
    final_loss = None
    for try_num in range(max_tries):                    # 54: restart while the fit error stays above 10
        fs.reset(trans_init)                            # 57-64 (zeros; trans at the target centroid)
        for stage, lr, n in (("coarse", 1e-1, epoch_coarse), ("fine", 1e-2, epoch_fine)):
            fs.new_optimizer()                          # 75 / 98: a fresh torch.optim.Adam per stage
            for _ in range(n):
                if use_graph:
                    graphs[stage].replay()
                else:
                    fs.iteration(lr, stage == "coarse")
        final_loss = float(fs.loss.item())             # loss of the last iteration's forward (107)
        if final_loss <= 10.0:
            break
    fs.forward()                                        # 120
    out = {"joints": (fs.joints * 1000.0).cpu().numpy(), "verts": (fs.verts0 * 1000.0).cpu().numpy(),
           "rot": fs.pose48[:, :3].cpu().numpy(), "pose": fs.pose48[:, 3:].cpu().numpy(),
           "shape": fs.shape.cpu().numpy(), "trans": fs.trans.cpu().numpy()}
    optimize_for_mano_param.last_loss = final_loss
    optimize_for_mano_param.last_tries = try_num + 1
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


# ---------------------------------------------------------------------------------------------------- sequence smoothing
def _sum_sq(x):
    return torch.sum(x ** 2)


def _temporal_smoothness(x, n_frames):
    """LossSmoothPoses(1, nFrames).poses (hand_utils.py:505-513): distance of every interior frame to the 3-frame mean,
    the mean held constant (detached)."""
    mean3 = x.detach().clone()
    mean3[1:-1] = (mean3[1:-1] + mean3[:-2] + mean3[2:]) / 3
    return _sum_sq(x[1:-1] - mean3[1:-1]) / (n_frames - 2)


def _keypoint_distance(est, target, n_frames):
    """LossKeypoints3D / LossAnchor .loss_func (hand_utils.py:472-496)."""
    nj = min(est.shape[1], target.shape[1], 21)
    return _sum_sq(est[:, :nj, :3] - target[:, :nj, :3]) / n_frames


def optimize_smooth_seq(params_cpu, mano_layer, use_smplx_arm=False, nimble_hand=False, img_res=224, mano_j_regressor=None,
                        total_iter_pose=1000, total_iter_cam=1000):
    """Same call and return value as `metro_modifications/hand_utils.py:540` (MANO and SMPL-X-arm branches): temporal smoothing of the per-frame
    fits.  Stage 1 optimises [rot, pose, shape, cam] with Adam(1e-3) on 1e-2 * (root-relative joints vs the input's) + 1e-1 *
    (temporal smoothness of the root-relative joints) until the running loss stops falling by 1e-5 — a host decision every
    iteration, as in the reference (592-595), so the loop is driven from Python: the LBS forward / backward of each iteration
    are `harp_lbs_mano_fwd/bwd` (through `ManoLayer`'s autograd function), the joint-space losses are a few tiny torch ops on
    (T,21,3) tensors.  Stage 2 optimises cam alone (Adam 1e-3 + ReduceLROnPlateau) on anchor + smoothness of the
    camera-relative root; the joints are constant there, so no skinning is needed.  Returns CPU tensors like the reference."""
    if nimble_hand:
        raise NotImplementedError("optimize_smooth_seq: the NIMBLE hand model needs third-party code and assets (SURVEY.md section 2)")
    dev = mano_layer.hc_coarse.device
    p = {k: torch.as_tensor(v, dtype=torch.float32).to(dev).clone() for k, v in params_cpu.items()}
    p["cam"] = p["cam"].unsqueeze(1)                                          # 547
    T = p["pose"].shape[0]
    target = p["joints"] - p["joints"][:, 0, :].unsqueeze(1)                  # 463-464
    free = [p["rot"], p["pose"], p["shape"], p["cam"]]
    for q in free:
        q.requires_grad_(True)
    if use_smplx_arm:  # hand_utils.py:586,620,639,663: the reference's default branch (utils/config_utils.py:7)
        skin = lambda: mano_layer(betas=p["shape"], global_orient=p["rot"], transl=p["trans"], right_hand_pose=p["pose"],
                                  return_type="mano", hc=mano_layer.hc_coarse)
    else:
        skin = lambda: mano_layer(torch.cat((p["rot"], p["pose"]), 1), p["shape"], p["trans"])
    opt = torch.optim.Adam(free, lr=1e-3)                                     # 585
    running = 999999.0
    for it in range(total_iter_pose):
        _, joints = skin()
        rel = joints - joints[:, 0, :].unsqueeze(1)
        loss = 1e-2 * _keypoint_distance(rel, target, T) + 1e-1 * _temporal_smoothness(rel, T)   # 566-569 (smooth_body: weight 0)
        now = loss.item()                                                    # the reference's per-iteration host sync
        if it > 0 and running - now < 0.00001:
            break
        running = (running + now) / 2.0
        opt.zero_grad()
        loss.backward()
        opt.step()
    with torch.no_grad():
        verts, joints = skin()
    p["joints"], p["verts"] = joints, verts                                  # 616-617
    focal = 1000.0 * (img_res / 224.0)
    root = joints[:, 0, :] / 1000.0

    def camera_relative_root(cam):
        t = torch.stack([cam[:, :, 1], cam[:, :, 2], 2 * focal / (img_res * cam[:, :, 0] + 1e-9)], dim=2)
        return t + root   # (T,1,3) + (T,3) broadcasts to (T,T,3) exactly as at hand_utils.py:643 / 665

    anchor = camera_relative_root(p["cam"]).detach()
    opt = torch.optim.Adam([p["cam"]], lr=1e-3)                               # 651
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=10)      # 652
    for it in range(total_iter_cam):
        crr = camera_relative_root(p["cam"])
        loss = 1e-2 * _keypoint_distance(crr, anchor, T) + 1e-2 * _temporal_smoothness(crr, T)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step(loss)
    p["cam"] = p["cam"].squeeze(1)
    return {k: v.detach().cpu() for k, v in p.items()}
