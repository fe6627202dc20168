"""Fused optimiser step: the body of `optimize_sequence.py:446-579` as one asynchronous launch sequence.

`HarpFitter` keeps every parameter, image and optimiser state resident in HBM (the reference keeps parameters on the CPU
and copies them every iteration: `utils/visualize.py:37-44,81,94`) and runs, per minibatch,

    LBS -> mesh build -> [silhouette fwd+L1+bwd] -> [light depth -> shade fwd+L1+bwd -> light depth bwd]
        -> mesh regularisers -> texture regularisers -> mesh build bwd -> LBS bwd -> (allreduce) -> Adam x2

with no host synchronisation: loss scalars stay on the device until the caller asks for them.
Loss weights, stage gating and optimiser grouping follow `optimize_sequence.py:253-310,411-422,507-515,567-573`.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import HarpRenderArgs, HarpShadeGrads, check, iptr, lptr, ptr, stream

# slots 1..5 are contiguous in harp_mesh_losses' output order
LOSS_NAMES = ("silhouette", "laplacian", "normal", "arap", "kps_anchor", "vert_disp_reg", "photo", "albedo", "normal_reg")
LOSS_WEIGHTS = {"silhouette": 7.0, "kps_anchor": 10.0, "vert_disp_reg": 2.0, "normal": 0.1, "laplacian": 4.0, "arap": 0.2,
                "photo": 1.0, "albedo": 0.5, "normal_reg": 0.1}  # optimize_sequence.py:411-422 (vgg excluded)
_SLOT = {n: i for i, n in enumerate(LOSS_NAMES)}


class _Flat:
    """Several named parameters as views of one flat fp32 buffer (+ grad, Adam moments) so that one fused Adam launch
    updates the whole optimiser group."""

    def __init__(self, named, device, grad_storage=None):
        self.names = [n for n, _ in named]
        sizes = [t.numel() for _, t in named]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        n = self.offsets[-1]
        self.data = torch.empty(n, device=device)
        self.grad = torch.zeros(n, device=device) if grad_storage is None else grad_storage[:n]
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device)
        self.p, self.g = {}, {}
        for (name, t), a, b in zip(named, self.offsets[:-1], self.offsets[1:]):
            self.data[a:b].copy_(t.detach().reshape(-1).to(device))
            self.p[name] = self.data[a:b].view(t.shape)
            self.g[name] = self.grad[a:b].view(t.shape)
        # {int32 step, two scratch words, float lr} on the device: the update is graph-capturable
        # and the learning rate, so a scheduler can change it under a captured graph
        self.state = torch.zeros(4, dtype=torch.int32, device=device)

    @property
    def steps(self):
        return int(self.state[0].item())

    def set_lr(self, lr):
        self.state[3:4].view(torch.float32).fill_(float(lr))

    def adam(self):
        check(_lib.lib().harp_adam_dev(ptr(self.data), ptr(self.grad), ptr(self.m), ptr(self.v), self.data.numel(), -1.0,
                                       0.9, 0.999, 1e-8, self.state.data_ptr(), stream()))


def shard_frames(total_frames, rank, world_size):
    """Static contiguous frame sharding: rank r owns [r T/R, (r+1) T/R); remainder to the first ranks."""
    base, rem = divmod(total_frames, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_shared(coarse, app, losses, world_size, group=None):
    """The step's ONE collective.  Per-frame rows (pose, cam) are rank-local under static frame sharding and are not
    exchanged; everything else is shared and summed: [verts_disps | shape | light row 0 | amb | texture | normal_map | losses]
    as a single flat fp32 buffer (about 6.3 MB).  Loss values come back averaged over ranks."""
    import torch.distributed as dist
    a0 = coarse.offsets[coarse.names.index("verts_disps")]
    shared_app = [app.g[n].reshape(-1) for n in app.names if n != "light_positions"]
    parts = [coarse.grad[a0:], app.g["light_positions"].reshape(-1)[:3]] + shared_app + [losses]
    flat = torch.cat([p.reshape(-1) for p in parts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for p in parts:
        n = p.numel()
        p.copy_(flat[o:o + n].view(p.shape))
        o += n
    losses /= world_size


def allreduce_shared_inplace(gbuf, start, stop, loss_slice, world_size, group=None):
    """The same exchange without staging copies, for gradients laid out in ONE buffer
    `[coarse: per-frame rows | verts_disps | shape] [losses] [app: amb | texture | normal_map | light (T,3)]`:
    a single in-place allreduce of `gbuf[start:stop]`, start = offset of verts_disps, stop = end of light row 0 — a span whose
    length is the same on every rank however the frames are sharded (the rank-local light rows 1..T-1 lie behind it)."""
    import torch.distributed as dist
    dist.all_reduce(gbuf[start:stop], op=dist.ReduceOp.SUM, group=group)
    gbuf[loss_slice] /= world_size


class _RawCudaBuffer:
    """A device allocation owned by the C library, exposed through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """The step's ONE collective as a kernel over NVLink peer memory (csrc/peer.cu, harp_peer_*): every rank's gradient buffer is
    an IPC-shareable allocation mapped by all ranks of the node; `allreduce` is a CUDA-graph-capturable launch on the current
    stream.  Handles are exchanged once through torch.distributed (any backend)."""

    def __init__(self, n_floats, rank, world_size, device, group=None):
        import torch.distributed as dist
        L = _lib.lib()
        self.h, data = C.c_void_p(), C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            check(L.harp_peer_create(C.byref(self.h), int(world_size), int(rank), int(n_floats) * 4, C.byref(data), handle))
            gathered = [None] * world_size
            dist.all_gather_object(gathered, bytes(handle), group=group)
            blob = (C.c_ubyte * (64 * world_size)).from_buffer_copy(b"".join(gathered))
            check(L.harp_peer_connect(self.h, blob))
        self._owner = _RawCudaBuffer(data.value, n_floats)
        self.buffer = torch.as_tensor(self._owner, device=device)
        dist.barrier(group=group)  # nobody launches against a peer that has not mapped everyone yet

    def allreduce(self, start, stop, scale_lo, scale_hi, scale):
        check(_lib.lib().harp_peer_allreduce(self.h, int(start), int(stop), int(scale_lo), int(scale_hi), float(scale), stream()))

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().harp_peer_destroy(self.h)
            self.h = None


class FrameStore:
    """the sequence's frames decoded ONCE and kept in HBM as uint8 — 4 B per pixel instead of the 20 B of the
    reference's float32 `y_true` / `y_sil` / `y_sil_eroded` tensors (ImagesDataset + load_img, utils/data_util.py:11-51, decode,
    divide and erode every frame on the host again in every epoch).  rgb (T,S,S,3) uint8, mask (T,S,S) uint8 (0 / 255),
    init_joints (T,21,3) mm.  Pass it to `HarpFitter(frames=...)`: every step converts (/255) and erodes its minibatch on the
    device (`harp_ingest_u8`, bit-identical to `load_img`)."""

    def __init__(self, rgb_u8, mask_u8, init_joints, device="cuda"):
        self.rgb = torch.as_tensor(rgb_u8).to(device=device, dtype=torch.uint8).contiguous()
        self.mask = torch.as_tensor(mask_u8).to(device=device, dtype=torch.uint8).contiguous()
        self.init_joints = torch.as_tensor(init_joints).to(device=device, dtype=torch.float32).contiguous()
        if self.rgb.dim() != 4 or self.rgb.shape[-1] != 3 or self.mask.shape != self.rgb.shape[:3]:
            raise ValueError("FrameStore: rgb (T,S,S,3) uint8 and mask (T,S,S) uint8 expected")

    def __len__(self):
        return self.rgb.shape[0]

    @property
    def nbytes(self):
        return self.rgb.numel() + self.mask.numel() + 4 * self.init_joints.numel()

    def batch(self, fid=None, frame_range=None):
        """-> dict(rgb, mask, init_joints) of a minibatch (views when a contiguous `frame_range` is promised)."""
        if frame_range is not None:
            a, b = frame_range
            return dict(rgb=self.rgb[a:b], mask=self.mask[a:b], init_joints=self.init_joints[a:b])
        return dict(rgb=self.rgb.index_select(0, fid), mask=self.mask.index_select(0, fid),
                    init_joints=self.init_joints.index_select(0, fid))


class HarpFitter:
    """params: dict with the reference's keys (`optimize_sequence.py:181-250`): pose (T,45), rot (T,3), trans (T,3),
    cam (T,3), shape (10,), verts_disps (V,1), texture / normal_map (1,H,W,3), light_positions (T,3), amb_ratio ().
    frames: y_true (T,S,S,3), y_sil (T,S,S), y_sil_eroded (T,S,S), init_joints (T,21,3) [mm] — device resident; or a
    `FrameStore` (uint8 frames + masks, converted and eroded per minibatch on the device).
    ref_verts: (1,V,3) ARAP reference (`optimize_sequence.py:429-435`); uv_mask (H,W)."""

    def __init__(self, hc, params, frames, ref_verts, uv_mask, img_size, focal_length, self_shadow=True,
                 lr_coarse=1e-3, lr_app=1e-2, world_size=1, process_group=None, opt_arm_pose=False, exchange="nccl", rank=None):
        """exchange (world_size > 1): 'nccl' = one in-place NCCL allreduce launched after the step (outside a captured graph);
        'peer' = the device-side exchange over NVLink peer memory (PeerExchange), part of the captured graph together with Adam."""
        dev = hc.device
        self.hc, self.S, self.focal, self.self_shadow = hc, int(img_size), float(focal_length), bool(self_shadow)
        self._lr = {"coarse": float(lr_coarse), "app": float(lr_app)}
        self.faces_per_pixel = 50  # silhouette rasteriser's K (optimize_sequence.py:427); settable for tests of the top-K path
        self.world_size, self.pg = int(world_size), process_group
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        # optimiser groups (get_optimizers, optimize_sequence.py:253-310, default `harp` branch).  With the SMPL-X arm model
        # (`use_arm`) and `opt_arm_pose`, wrist_pose and rot join the coarse group at the same lr (279-283); they are
        # placed BEFORE verts_disps so that everything from verts_disps on is the shared (all-reduced) block.
        self.arm = hc.J == 55
        T_ = params["pose"].shape[0]
        self.wrist = f32(params["wrist_pose"]) if self.arm and "wrist_pose" in params else torch.zeros(T_, 3, device=dev)
        groups = [("pose", f32(params["pose"])), ("cam", f32(params["cam"]))]
        self.opt_arm_pose = bool(opt_arm_pose and self.arm)
        if self.opt_arm_pose:
            groups += [("wrist_pose", self.wrist), ("rot", f32(params["rot"]))]
        groups += [("verts_disps", f32(params["verts_disps"])), ("shape", f32(params["shape"]))]
        # the rank-local light rows go last so that everything shared is one contiguous span (see allreduce_shared_inplace)
        app_groups = [("amb_ratio", f32(params["amb_ratio"]).view(1)), ("texture", f32(params["texture"])),
                      ("normal_map", f32(params["normal_map"])), ("light_positions", f32(params["light_positions"]))]
        # all gradients and the loss vector live in one buffer: one fill per step, one in-place allreduce of its shared span
        nc, na, nl = sum(t.numel() for _, t in groups), sum(t.numel() for _, t in app_groups), len(LOSS_NAMES)
        self.exchange_mode, self.peer = (exchange if int(world_size) > 1 else "none"), None
        if self.exchange_mode == "peer":
            import torch.distributed as dist
            self.peer = PeerExchange(nc + nl + na, dist.get_rank(process_group) if rank is None else rank, int(world_size), dev,
                                     group=process_group)
            self._gbuf = self.peer.buffer
            self._gbuf.zero_()
        else:
            self._gbuf = torch.zeros(nc + nl + na, device=dev)
        self.coarse = _Flat(groups, dev, grad_storage=self._gbuf[:nc])
        self.app = _Flat(app_groups, dev, grad_storage=self._gbuf[nc + nl:])
        self.losses = self._gbuf[nc:nc + nl]
        self._loss_slice = slice(nc, nc + nl)
        self._shared_start = self.coarse.offsets[self.coarse.names.index("verts_disps")]
        self._shared_stop = nc + nl + self.app.offsets[self.app.names.index("light_positions")] + 3
        self.coarse.set_lr(self._lr["coarse"])
        self.app.set_lr(self._lr["app"])
        self.rot, self.trans = f32(params["rot"]), f32(params["trans"])
        if self.opt_arm_pose:
            self.rot, self.wrist = self.coarse.p["rot"], self.coarse.p["wrist_pose"]
        self.store = frames if isinstance(frames, FrameStore) else None
        self.frames = {"init_joints": self.store.init_joints} if self.store is not None else {k: f32(v) for k, v in frames.items()}
        self.ref_verts = f32(ref_verts).view(-1, 3)
        self.uv_mask = f32(uv_mask)
        self.T = self.rot.shape[0]
        # everything of `params` that is not optimised here (mesh_faces, faces_uvs, verts_uvs, uv_mask, init_joints, verts_rgb,
        # wrist_pose of a frozen arm, ...) is kept so that a checkpoint written from the fitter has every key the reference's
        # load_result / prepare_mesh / prepare_materials index (utils/visualize.py:48,85-86)
        own = {n for n, _ in groups} | {n for n, _ in app_groups} | {"amb_ratio", "rot", "trans"}
        self.extra_params = {k: v for k, v in params.items() if k not in own}
        self._scratch = {}
        self.last = {}

    # learning rates live on the device (read by harp_adam_dev), so assigning them also works under a captured graph —
    # what the reference's two ReduceLROnPlateau schedulers need (optimize_sequence.py:312-321, 601-603)
    @property
    def lr_coarse(self):
        return self._lr["coarse"]

    @lr_coarse.setter
    def lr_coarse(self, v):
        self._lr["coarse"] = float(v)
        self.coarse.set_lr(v)

    @property
    def lr_app(self):
        return self._lr["app"]

    @lr_app.setter
    def lr_app(self, v):
        self._lr["app"] = float(v)
        self.app.set_lr(v)

    # -- helpers --------------------------------------------------------------------------------------------------------
    def params(self):
        out = dict(self.coarse.p)
        out.update(self.app.p)
        out["amb_ratio"] = self.app.p["amb_ratio"].view(())
        out["rot"], out["trans"] = self.rot, self.trans
        if self.arm:
            out["wrist_pose"] = self.wrist
        for k, v in self.extra_params.items():
            out.setdefault(k, v)
        return out

    def _buf(self, name, shape, zero=True):
        t = self._scratch.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, device=self.hc.device)
            self._scratch[name] = t
        if zero:
            t.zero_()
        return t

    # -- one minibatch ----------------------------------------------------------------------------------------------------
    def step(self, fid, off_albedo, off_normal, stage="both", update=True, batch=None, frame_range=None, exchange=True,
             batch_u8=None):
        """fid: (B,) long device tensor of frame ids owned by this rank.  stage: 'coarse' | 'both' | 'app'
        (`optimize_sequence.py:507-515`).  batch: optional dict overriding self.frames rows (host->device ingest path).
        frame_range=(a,b): promise that fid == arange(a,b), so image rows are used in place instead of gathered.
        batch_u8: dict(rgb (B,S,S,3) uint8, mask (B,S,S) uint8) device tensors as decoded from the image files; converted
        (and the mask eroded) on the device like `load_img` (utils/data_util.py:11-30).
        Returns the device tensor of unweighted losses (LOSS_NAMES order); gradients are in self.coarse.grad / self.app.grad."""
        hc, L, S, focal = self.hc, _lib.lib(), self.S, self.focal
        st = stream()
        B = int(fid.shape[0])
        R = self.world_size
        hc.reserve(B, S)
        P, A = self.coarse.p, self.app.p
        if batch is None and batch_u8 is None and self.store is not None:
            batch_u8 = self.store.batch(fid, frame_range)
        if batch_u8 is not None:
            from .ops import ingest_u8
            fr = ingest_u8(batch_u8["rgb"], batch_u8["mask"],
                           out=dict(y_true=self._buf("u8_y_true", (B, S, S, 3), zero=False), y_sil=self._buf("u8_y_sil", (B, S, S), zero=False),
                                    y_sil_eroded=self._buf("u8_y_er", (B, S, S), zero=False)))
            fr["init_joints"] = batch_u8["init_joints"] if "init_joints" in batch_u8 else self.frames["init_joints"].index_select(0, fid)
        elif batch is not None:
            fr = batch
        elif frame_range is not None:
            fr = {k: v[frame_range[0]:frame_range[1]] for k, v in self.frames.items()}
        else:
            fr = {k: v.index_select(0, fid) for k, v in self.frames.items()}
        self._gbuf.zero_()  # losses, coarse and appearance gradients
        ws = self._buf("workspace", (int(L.harp_step_workspace_floats(hc.h, B)),), zero=False)
        alpha = self._buf("alpha", (B, S, S), zero=False)
        rgb = self._buf("rgb", (B, S, S, 3), zero=False)
        d_rot, d_trans, d_wrist = (self._buf(n, (B, 3), zero=False) for n in ("d_rot", "d_trans", "d_wrist"))
        tex, nm = A["texture"], A["normal_map"]
        G, GA = self.coarse.g, self.app.g
        disps = P["verts_disps"].view(-1)
        a = _lib.HarpStepArgs(
            B=B, S=S, stage={"coarse": 1, "app": 2, "both": 3}[stage], self_shadow=int(self.self_shadow),
            faces_per_pixel=int(self.faces_per_pixel), texH=tex.shape[-3], texW=tex.shape[-2], focal=focal, sigma=1e-7,
            frame_scale=1.0 / R, weights=(C.c_float * 9)(*[LOSS_WEIGHTS[n] for n in LOSS_NAMES]),
            fid=lptr(fid), pose=ptr(P["pose"]), rot=ptr(self.rot), trans=ptr(self.trans), cam=ptr(P["cam"]),
            wrist_pose=ptr(self.wrist) if self.arm else None, shape=ptr(P["shape"]), verts_disps=ptr(disps), texture=ptr(tex),
            normal_map=ptr(nm), light_position=ptr(A["light_positions"]), amb_ratio=ptr(A["amb_ratio"]),
            y_true=ptr(fr["y_true"]), y_sil=ptr(fr["y_sil"]), y_sil_eroded=ptr(fr["y_sil_eroded"]), init_joints=ptr(fr["init_joints"]),
            ref_verts=ptr(self.ref_verts), uv_mask=ptr(self.uv_mask), off_albedo=iptr(off_albedo), off_normal=iptr(off_normal),
            losses=ptr(self.losses), alpha=ptr(alpha), rgb=ptr(rgb), g_pose=ptr(G["pose"]), g_cam=ptr(G["cam"]),
            g_rot=ptr(G["rot"]) if self.opt_arm_pose else None, g_wrist_pose=ptr(G["wrist_pose"]) if self.opt_arm_pose else None,
            g_shape=ptr(G["shape"]), g_verts_disps=ptr(G["verts_disps"].view(-1)), g_texture=ptr(GA["texture"]),
            g_normal_map=ptr(GA["normal_map"]), g_light_position=ptr(GA["light_positions"]), g_amb_ratio=ptr(GA["amb_ratio"]),
            d_rot=ptr(d_rot), d_trans=ptr(d_trans), d_wrist=ptr(d_wrist) if self.arm else None, workspace=ptr(ws))
        check(L.harp_step_fused(hc.h, C.byref(a), st))
        views = [C.c_void_p() for _ in range(4)]
        check(L.harp_step_workspace_views(hc.h, B, ptr(ws), *[C.byref(v) for v in views]))
        off = [(v.value - ws.data_ptr()) // 4 for v in views]
        NJ = hc.n_joints_out
        self.last = dict(d_rot=d_rot, d_trans=d_trans, alpha=alpha, rgb=rgb, d_wrist=d_wrist if self.arm else None,
                         verts0=ws[off[0]:off[0] + B * hc.V0 * 3].view(B, hc.V0, 3), joints=ws[off[1]:off[1] + B * NJ * 3].view(B, NJ, 3),
                         verts=ws[off[2]:off[2] + B * hc.V * 3].view(B, hc.V, 3), normals=ws[off[3]:off[3] + B * hc.V * 3].view(B, hc.V, 3))
        # ---- data-parallel exchange: ONE exchange of the shared-parameter gradients
        if R > 1 and exchange:
            self._allreduce_shared()
        if update:
            self._update(stage)
        return self.losses

    def _update(self, stage):
        if stage in ("coarse", "both"):
            self.coarse.adam()
        if stage in ("both", "app"):
            self.app.adam()

    # -- CUDA-graph replay ------------------------------------------------------------------------------------------------
    def capture(self, fid, off_albedo, off_normal, stage="both", frame_range=None, batch=None, warmup=2, keep_profile=False,
                batch_u8=None):
        """Record one optimiser step (all C-ABI launches + the small torch glue) into a CUDA graph.  The ~120 launches of
        a step then replay back to back; inputs are read from the same device buffers every replay (fid, offsets, the
        resident frames or the staging `batch`).  Returns a callable; each call runs one step (update=True)."""
        multi = self.world_size > 1 and self.exchange_mode != "peer"  # NCCL: the allreduce and the two Adam launches stay
        # outside the graph; the peer-memory exchange is a kernel node like any other
        kw = dict(stage=stage, frame_range=frame_range, batch=batch, batch_u8=batch_u8)
        if not keep_profile:
            self.hc.profile(False)
        side = torch.cuda.Stream(device=self.hc.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(fid, off_albedo, off_normal, update=True, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if keep_profile:
            self.hc.profile_read()  # drop the warm-up event pairs; the captured step re-creates one pair per kernel
        graph = torch.cuda.CUDAGraph()
        # captured on a high-priority stream: kernel nodes inherit it, so the critical path (camera bins -> pixel pass ->
        # shading -> geometry backward) is scheduled ahead of the two side branches (default / least priority)
        if getattr(self, "_capture_stream", None) is None:
            self._capture_stream = torch.cuda.Stream(device=self.hc.device, priority=-1)
        with torch.cuda.graph(graph, stream=self._capture_stream):
            self.step(fid, off_albedo, off_normal, update=not multi, exchange=not multi, **kw)
        self._set_graph(graph)
        if not multi:
            return graph.replay

        def replay_and_exchange():
            graph.replay()
            self._allreduce_shared()
            self._update(stage)
        return replay_and_exchange

    def _set_graph(self, graph):
        """Keep every captured graph and pin the context's scratch for as long as they live (a later harp_ctx_reserve with
        a larger batch or another image size must not free memory a graph still replays into)."""
        if getattr(self, "_graphs", None) is None:
            self._graphs = []
        self._graphs.append(graph)
        self._graph = graph
        self.hc.pin(+1)

    def release_graph(self):
        """Drop all captured graphs of this fitter and unpin the context."""
        n = len(getattr(self, "_graphs", None) or [])
        self._graphs, self._graph = [], None
        if n:
            self.hc.pin(-n)

    def __del__(self):
        try:
            self.release_graph()
        except Exception:
            pass

    def close(self):
        """Release captured graphs and the peer-memory exchange (all ranks together: peers unmap each other's buffers)."""
        self.release_graph()
        if self.peer is not None:
            torch.cuda.synchronize()
            self.peer.close()
            self.peer = None

    def _allreduce_shared(self):
        if self.peer is not None:
            self.peer.allreduce(self._shared_start, self._shared_stop, self._loss_slice.start, self._loss_slice.stop, 1.0 / self.world_size)
        else:
            allreduce_shared_inplace(self._gbuf, self._shared_start, self._shared_stop, self._loss_slice, self.world_size, self.pg)

    def loss_dict(self):
        """One device->host read of the loss vector (the reference does >= 10 `.cpu()` syncs per iteration, 556-561)."""
        vals = self.losses.detach().cpu().tolist()
        cnt = self.hc.counters()  # this call already synchronises: check the only two conditions that make results inexact
        if cnt["bin_overflow"] or cnt["sil_k_overflow_pixels"]:
            import warnings
            warnings.warn(f"harp_b200: results are not exact for this framing: {cnt} (a hand zoomed far past the image, or faces "
                          "crossing the camera plane, overflowed the tile lists / the exact top-K path; see DESIGN.md section 2)")
        self.last_counters = cnt
        d = dict(zip(LOSS_NAMES, vals))
        d["total"] = sum(LOSS_WEIGHTS[k] * v for k, v in d.items())
        return d
