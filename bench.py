#!/usr/bin/env python
"""Headline benchmark: optimiser frames/s (forward + backward + Adam [+ allreduce]) of the HARP render-and-fit step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl harp|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Workload (BASELINE.json configs[1]): a 64-frame sequence at 512x512 per GPU, 778-vert MANO subdivided to 3093 verts /
6152 faces, every loss of `optimize_sequence.py:517-553` except VGG, stage "shape and appearance".
One step = one minibatch holding the rank's whole 64-frame shard.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, SIZE = 64, 512
METRIC, UNIT = "optimizer_frames_per_sec_fwd_bwd_512x512_MANO", "frames/s"


def config(n_gpus, frames, size, model="hand"):
    mesh = "MANO 778->3093 verts / 6152 faces" if model == "hand" else "SMPL-X arm 1026->4083 verts / 8128 faces"
    return {"workload": f"{frames}-frame seq @{size}x{size} per GPU, {mesh}, soft silhouette (K=50) + "
                        "shadowed PBR render + photometric + kps + Laplacian + normal-consistency + ARAP + texture regs, "
                        "fwd+bwd+Adam (VGG term excluded)",
            "frames_per_gpu": frames, "img_size": size, "global_frames": frames * n_gpus,
            "parallelism": f"dp{n_gpus} (static round-robin frame sharding, one exchange of shared-parameter grads per step)",
            "l2_policy": f"per-step image inputs ({20 * size * size * frames / 1e6:.0f} MB/GPU) exceed the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(steps, warmup, frames, size, device="cpu"):
    """The reference's CPU implementation of the path = the oracle (reference Python on the restated PyTorch3D shim is
    pinned to it bit-for-bit, tests/test_oracle_golden.py).  Each step is a bounded sample: `frames` frames.
    device='cuda': the same eager PyTorch program with every tensor on the GPU — the stand-in for "the reference's 1-GPU
    PyTorch3D path" (PyTorch3D's own CUDA rasteriser cannot be installed here), reported as
    `gpu_eager_standin`, never as the product."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harp_oracle as O
    from harp_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device(device)
    mv = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}
    tmpl = synth.load_template("hand")
    pack = synth.make_mano_pack(seed=0, template=tmpl)
    model = {k: torch.tensor(v) for k, v in pack.items()}
    topo = O.Topology(pack["th_faces"], tmpl["faces_uvs"], tmpl["verts_uvs"], subdivide=True)
    seq = synth.make_sequence(FRAMES, seed=1, img_size=size)
    app = synth.make_appearance(seed=2)
    B, S = frames, size
    focal = 1000.0 * S / 224.0
    params = dict(pose=seq["pose"][:B], rot=seq["rot"][:B], trans=seq["trans"][:B], cam=seq["cam"][:B], shape=seq["shape"],
                  verts_disps=torch.zeros(topo.V, 1), texture=app["texture"], normal_map=app["normal_map"],
                  light_position=app["light_position"], amb_ratio=app["amb_ratio"])
    cfg = dict(img_size=S, focal_length=focal, stage="both", self_shadow=True)
    with torch.no_grad():
        tp = dict(params)
        tp["cam"] = params["cam"] + torch.tensor([0.1, 0.003, -0.002])
        verts_mm, joints_mm = O.mano_forward(model, torch.cat((tp["rot"], tp["pose"]), 1), tp["shape"].repeat(B, 1), tp["trans"])
        verts = O.build_mesh(verts_mm / 1000.0, topo, None)
        y_sil = (O.render_silhouette(verts, topo, tp["cam"], focal, S) >= 0.5).float()
        y_true = O.render_rgb_shadow(verts, topo, tp["cam"], tp["light_position"][None].repeat(B, 1), tp["amb_ratio"],
                                     synth.make_appearance(seed=5)["texture"], tp["normal_map"], focal, S).clamp(0, 1)
    batch = dict(y_true=y_true, y_sil=y_sil, y_sil_eroded=synth.erode_mask(y_sil), init_joints=joints_mm.detach(),
                 ref_verts=verts[:1].detach(), uv_mask=torch.tensor(tmpl["uv_mask"]),
                 off_albedo=synth.make_reg_offsets(3, 1.0), off_normal=synth.make_reg_offsets(4, 2.0))
    if dev.type == "cuda":
        model, params, batch = mv(model), mv(params), mv(batch)
        for k in ("faces", "faces0", "edges0", "faces_uvs", "verts_uvs"):
            if torch.is_tensor(getattr(topo, k, None)):
                setattr(topo, k, getattr(topo, k).to(dev))
    import contextlib
    times = []
    # on the GPU the oracle's own factory calls (torch.zeros / eye / tensor ...) must default to the device too
    with (torch.device(dev) if dev.type == "cuda" else contextlib.nullcontext()):
        for i in range(warmup + steps):
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            O.step(model, topo, params, batch, cfg)
            if dev.type == "cuda":
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=B / sec, ms_per_step=sec * 1e3, cores=cores,
                sample=f"{B} frame(s) @{S}x{S} per step, full loss set, fwd+bwd via autograd, {steps} steps after {warmup} warm-up")


def dropin_leg(sc, hc_model, B, S, steps, warmup):
    """The reference's loop body (optimize_sequence.py:446-579) written against the Level-1 drop-in modules:
    prepare_mesh -> silhouette renderer -> shadow renderer -> losses -> autograd backward -> two torch.optim.Adam steps.
    This is what a HARP maintainer gets by swapping imports; HarpFitter.step is the fused Level-3 path."""
    from harp_b200 import renderer_helper, synth
    from harp_b200.manolayer import ManoLayer
    from harp_b200.structures import Meshes
    from harp_b200.visualize import prepare_mesh, prepare_materials, render_image, render_image_with_RT
    from harp_b200.loss import arap_loss, kps_loss, albedo_reg, normal_reg, mesh_laplacian_smoothing, mesh_normal_consistency
    from harp_b200.step import LOSS_WEIGHTS as W
    dev = "cuda"
    tmpl = synth.load_template("hand")
    layer = ManoLayer(synth.make_mano_pack(seed=0, template=tmpl), faces_uvs=tmpl["faces_uvs"], verts_uvs=tmpl["verts_uvs"], device=dev)
    P = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in sc["params"].items()}
    params = dict(P)
    params["mesh_faces"] = layer.th_faces
    params["verts_uvs"] = torch.tensor(tmpl["verts_uvs"])[None]
    params["faces_uvs"] = torch.tensor(tmpl["faces_uvs"]).long()[None]
    fr = sc["frames"]
    focal = sc["focal"]
    coarse = torch.optim.Adam([P[k] for k in ("pose", "cam", "verts_disps", "shape")], lr=1e-3)
    app = torch.optim.Adam([P[k] for k in ("amb_ratio", "texture", "normal_map", "light_positions")], lr=1e-2)
    fid = torch.arange(B)
    configs = {"model_type": "harp", "img_size": S, "focal_length": focal, "use_arm": False}
    l1 = torch.nn.L1Loss()
    ref_mesh = None

    def body():
        nonlocal ref_mesh
        light_positions = params["light_positions"][0].repeat(B, 1)
        _, sil_r, _ = renderer_helper.get_renderers(image_size=S, light_posi=light_positions, silh_sigma=1e-7, silh_gamma=1e-1,
                                                    silh_faces_per_pixel=50, device=dev)
        joints, verts, faces, textures = prepare_mesh(params, fid, layer, False, object(), global_pose=False, configs=configs,
                                                      shared_texture=True, device=dev, use_arm=False)
        meshes = Meshes(verts, faces, textures)
        cam = params["cam"][fid]
        mats = prepare_materials(params, B, device=dev)
        y_sil_pred = render_image(meshes, cam, B, sil_r, S, focal, silhouette=True, device=dev)
        lR, lT, cR, cT = renderer_helper.process_info_for_shadow(cam, light_positions, verts.mean(1), image_size=S, focal_length=focal, device=dev)
        sh_r = renderer_helper.get_shadow_renderers(image_size=S, light_posi=light_positions, amb_ratio=torch.sigmoid(params["amb_ratio"]), device=dev)
        y_pred = render_image_with_RT(meshes, lT, lR, cT, cR, B, sh_r, S, focal, silhouette=False, materials_properties=mats, device=dev)
        if ref_mesh is None:
            ref_mesh = Meshes(sc["ref_verts"].to(dev), faces[:1])
        me = fr["y_sil_eroded"].unsqueeze(-1)
        loss = (W["silhouette"] * l1(fr["y_sil"], y_sil_pred) + W["kps_anchor"] * kps_loss(fr["init_joints"], joints, device=dev) +
                W["vert_disp_reg"] * torch.sum(params["verts_disps"] ** 2.0) + W["laplacian"] * mesh_laplacian_smoothing(meshes) +
                W["normal"] * mesh_normal_consistency(meshes) + W["arap"] * arap_loss(meshes, ref_mesh) +
                W["photo"] * l1(fr["y_true"] * me, y_pred * me) +
                W["albedo"] * albedo_reg(params["texture"], uv_mask=sc["uv_mask"].to(dev), std=1.0) +
                W["normal_reg"] * normal_reg(params["normal_map"], uv_mask=sc["uv_mask"].to(dev)))
        coarse.zero_grad(); app.zero_grad()
        loss.backward()
        coarse.step(); app.step()
        return loss

    for _ in range(warmup):
        body()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = body()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "what": "optimize_sequence.py:446-579 on the Level-1 drop-in modules (separate silhouette / shadow renderer calls, "
                    "torch.autograd, torch.optim.Adam, eager launches)", "loss": float(loss)}


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="harp", choices=["harp", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--model", default="hand", choices=["hand", "arm"], help="arm = SMPL-X arm sub-mesh (BASELINE.json configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--exchange", default=os.environ.get("HARP_EXCHANGE", "peer"), choices=["peer", "nccl"],
                    help="N > 1: peer = exchange kernel over NVLink peer memory inside the step graph (default); nccl = one NCCL "
                         "allreduce after the graph")
    ap.add_argument("--no-standin", action="store_true", help="skip the GPU-eager stand-in and drop-in legs (N = 1 only)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    args.warmup = max(args.warmup, 3) if args.impl == "harp" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return
        # four frames per step (~1.7 s on 16 cores; the CPU path is fastest per frame around this batch): honour K and W,
        # bounded so the arm ends within a few minutes
        k, w = max(1, min(args.steps, 60)), max(0, min(args.warmup, 5))
        r = cpu_reference(k, w, frames=4, size=args.size)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": k,
                "warmup": w, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": dict(config(args.gpus, args.frames, args.size), reference_sample_frames_per_step=4,
                               note="the CPU arm steps a bounded 4-frame sample of the 64-frame workload; value is frames/s"),
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # anything a library writes to fd 1 (NCCL prints its version banner there) goes to stderr: stdout carries ONE JSON line
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch.distributed as dist
    from harp_b200 import _lib, scene
    from harp_b200.step import HarpFitter, LOSS_NAMES
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl harp needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.frames, args.size
    hc, _, _ = scene.make_context(device=dev, model=args.model)
    # static round-robin sharding of the world*B-frame sequence: rank r owns frames r, r + R, ... (every rank sees the whole
    # motion cycle, so the content-dependent raster work is balanced across ranks)
    sc = scene.make_scene(hc, B, S, seed=1, frame_offset=rank, total_frames=world * B, frame_stride=world)
    exchange = args.exchange if world > 1 else "none"
    if exchange == "peer":  # all ranks must agree: fall back to NCCL together if the IPC mapping fails anywhere
        ok = torch.ones(1, device=dev)
        try:
            ft = HarpFitter(hc, sc["params"], sc["frames"], sc["ref_verts"], sc["uv_mask"], S, sc["focal"], world_size=world, exchange="peer")
        except Exception as e:  # pragma: no cover
            print(f"[bench] peer-memory exchange unavailable on rank {rank}: {e}", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            exchange = "nccl"
    if exchange != "peer":
        ft = HarpFitter(hc, sc["params"], sc["frames"], sc["ref_verts"], sc["uv_mask"], S, sc["focal"], world_size=world)
    fid = torch.arange(B, device=dev)
    oa, on = sc["offsets"]["albedo"], sc["offsets"]["normal"]

    def run_step():
        return ft.step(fid, oa, on, stage="both", update=True, frame_range=(0, B))

    n_before = _lib.lib().harp_launch_count()
    run_step()
    launches_per_step = _lib.lib().harp_launch_count() - n_before
    use_graph = not args.eager
    if use_graph:
        # the step is captured with the event brackets of the two pixel kernels inside (event-record graph nodes), so every
        # replay re-records them and the kernel durations are read from the timed region itself
        hc.profile(True)
        try:
            replay = ft.capture(fid, oa, on, stage="both", frame_range=(0, B), keep_profile=True)
            run_step = replay
        except Exception as e:  # pragma: no cover - falls back to eager launches
            print(f"[bench] CUDA graph capture failed ({e}); timing eager launches", file=sys.stderr)
            use_graph = False

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        run_step()
    sync()
    hc.counters()
    if not use_graph:
        hc.profile(True)
        hc.profile_read()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = _lib.lib().harp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    torch.cuda.nvtx.range_push("harp_timed")  # ncu --nvtx --nvtx-include "harp_timed/" isolates the timed region
    for _ in range(args.steps):
        run_step()
    torch.cuda.nvtx.range_pop()
    e1.record()
    sync()
    launches = launches_per_step * args.steps if use_graph else _lib.lib().harp_launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    per_rank_ms = [float(ms) / args.steps]
    if world > 1:
        allms = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(allms, ms)
        per_rank_ms = [float(x) / args.steps for x in allms]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clk = clocks.stop() if rank == 0 else None
    prof = hc.profile_read()
    hc.profile(False)
    cnt = hc.counters()
    cnt["deferred_pixels_last_step"] = hc.deferred_pixels()  # exact top-K fallback (slower path), of B*S*S pixels
    losses = ft.loss_dict()

    def e2e_leg(host, make_stage, make_run, n):
        """Double-buffered: uploads on a copy stream, step i+1's upload overlaps step i's kernels; the loss vector of
        step i is read back (pinned, async) while step i+1 runs.  Returns wall seconds for n steps."""
        stages = [make_stage() for _ in range(2)]
        runs = [make_run(s) for s in stages]
        copy_stream, main = torch.cuda.Stream(device=dev), torch.cuda.current_stream()
        loss_host = [torch.empty(len(LOSS_NAMES), pin_memory=True) for _ in range(2)]
        copied, done, ready = ([torch.cuda.Event() for _ in range(2)] for _ in range(3))

        def upload(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[s])  # the step that last read this staging buffer has finished
                for k, v in host.items():
                    stages[s][k].copy_(v, non_blocking=True)
                copied[s].record(copy_stream)

        def loop(k):
            seen = []
            upload(0)
            for i in range(k):
                s = i % 2
                if i + 1 < k:
                    upload(i + 1)
                main.wait_event(copied[s])
                runs[s]()
                done[s].record(main)
                loss_host[s].copy_(ft.losses, non_blocking=True)
                ready[s].record(main)
                if i > 0:
                    ready[1 - s].synchronize()
                    seen.append(float(loss_host[1 - s][0]))
            ready[(k - 1) % 2].synchronize()
            seen.append(float(loss_host[(k - 1) % 2][0]))
            return seen

        for ev in done:
            ev.record(main)
        loop(2)
        sync()
        t0 = time.perf_counter()
        loop(n)
        sync()
        sec = torch.tensor([max(time.perf_counter() - t0, 1e-9)], device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec)

    # ---- end to end through the public call with HOST image buffers (pinned): every step copies its frames host->device
    # and reads the loss vector back.  Leg 1 (`e2e_float32_host`): fp32 y_true / y_sil / y_sil_eroded, the dtypes the reference's
    # DataLoader yields (20 B/pixel).  Leg 2 (`e2e`): the uint8 frames + masks the image files decode to (4 B/pixel), converted
    # and eroded on the device (harp_ingest_u8): the plugin's ingest API.
    from harp_b200.hostmem import gpu_numa_affinity
    with gpu_numa_affinity(dev) as host_affinity:  # pinned staging pages on the GPU's NUMA node
        host = {k: ft.frames[k].cpu().pin_memory() for k in ("y_true", "y_sil", "y_sil_eroded")}
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = len(LOSS_NAMES) * 4
        k_e2e = 0 if args.no_e2e else max(3, min(args.steps, 50))  # (the first upload of a leg is not overlapped: amortised)
        e2e_s = e2e_u8_s = 1e-9
        e2e_res = None
        h2d_u8 = 0
        if k_e2e:
            def stage_f32():
                s = {k: torch.empty_like(ft.frames[k]) for k in host}
                s["init_joints"] = ft.frames["init_joints"]
                return s

            def run_f32(s):
                if use_graph:
                    return ft.capture(fid, oa, on, stage="both", batch=s, warmup=1)
                return lambda: ft.step(fid, oa, on, stage="both", update=True, batch=s)

            e2e_s = e2e_leg(host, stage_f32, run_f32, k_e2e)
            host_u8 = {"rgb": (ft.frames["y_true"] * 255).round().clamp(0, 255).to(torch.uint8).cpu().pin_memory(),
                       "mask": (ft.frames["y_sil"] * 255).round().to(torch.uint8).cpu().pin_memory()}
            h2d_u8 = sum(v.numel() for v in host_u8.values())

            def stage_u8():
                return {k: torch.empty(v.shape, dtype=torch.uint8, device=dev) for k, v in host_u8.items()}

            def run_u8(s):
                if use_graph:
                    return ft.capture(fid, oa, on, stage="both", batch_u8=s, warmup=1)
                return lambda: ft.step(fid, oa, on, stage="both", update=True, batch_u8=s)

            e2e_u8_s = e2e_leg(host_u8, stage_u8, run_u8, k_e2e)

        # Epoch-level end to end (how HARP uses the path: hundreds of epochs over the same frames, optimize_sequence.py:
        # 440-446): the sequence is uploaded ONCE from the pinned host buffers into the fitter's resident frame tensors, then
        # every step runs on resident frames and returns its loss vector to the host.
        if k_e2e:
            k_res = max(k_e2e, min(args.steps, 300))
            lh = torch.empty(k_res, len(LOSS_NAMES), pin_memory=True)
            sync()
            t0 = time.perf_counter()
            for k, v in host.items():
                ft.frames[k].copy_(v, non_blocking=True)
            for i in range(k_res):
                run_step()
                lh[i].copy_(ft.losses, non_blocking=True)
            sync()
            sec = torch.tensor([max(time.perf_counter() - t0, 1e-9)], device=dev)
            if world > 1:
                dist.all_reduce(sec, op=dist.ReduceOp.MAX)
            e2e_res = {"value": world * B * k_res / float(sec), "unit": UNIT, "steps": k_res, "h2d_bytes_total": h2d,
                       "d2h_bytes_per_step": d2h,
                       "what": "frames uploaded once (fp32, pinned host -> device), then every step on resident frames with its loss "
                               "vector read back: the reference's epoch loop over a fixed sequence"}

    if rank == 0:
        value = world * B * args.steps / (ms_total * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        # dominant kernel: the longer of the two pixel kernels; algorithmic bytes per launch:
        #   shade fwd+bwd: read y_true 12 S^2 + eroded mask 4 S^2, write rgb 12 S^2  = 28 S^2 B/frame
        #   silhouette fwd+bwd: read y_sil 4 S^2, write alpha 4 S^2                   =  8 S^2 B/frame
        # with the fused pixel pass, kernel id 0 is the tile rasteriser (silhouette + visibility over the covered tiles; the
        # background pixels of both images and their loss terms are its and the empty-tile pass's) and id 1 is the deferred
        # shading of the compacted hit pixels
        hit = float((ft.last["alpha"] > 0.5).float().mean())
        alg = {"shade": 28.0 * S * S * B * hit, "silhouette": (8.0 + 28.0 * (1.0 - hit)) * S * S * B}
        names = {"silhouette": "tile_raster_kernel<silhouette fwd+L1+bwd, visibility, hit compaction>",
                 "shade": "shade_deferred_kernel<shading fwd+L1+bwd over compacted hit pixels>"}
        dom = max(prof, key=lambda n: prof[n][0])
        dms = prof[dom][0] / max(prof[dom][1], 1)
        achieved = alg[dom] / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
        traffic = None
        try:
            if args.model == "hand" and S == SIZE and B == FRAMES:  # the ncu capture is of the default workload
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config(world, B, S, args.model),
                "clocks": clk,
                # e2e = the plugin's ingest API (HarpFitter.step(batch_u8=...) -> harp_ingest_u8): host buffers are the decoded uint8
                # frames + masks as the image files hold them; the float conversion and the mask erosion of the reference's
                # load_img (utils/data_util.py:11-30) happen on the device, bit-identically (tests/test_io_fixtures.py).
                # e2e_float32_host = the same step fed the reference DataLoader's float32 tensors (20 B/pixel): PCIe-bound.
                "e2e": {"value": world * B * k_e2e / e2e_u8_s, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                        "steps": k_e2e, "host_dtype": "uint8 frames + masks (what load_img decodes); /255 and 3x3x2 erosion on the device",
                        "host_affinity": host_affinity},
                "e2e_float32_host": {"value": world * B * k_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                     "steps": k_e2e, "host_dtype": "fp32 y_true / y_sil / y_sil_eroded (the reference DataLoader's tensors)"},
                "e2e_resident_epochs": e2e_res,
                "gpu_launches": int(launches), "launch_mode": "cuda_graph_replay" if use_graph else "eager",
                "exchange": {"none": "single GPU", "peer": "reduce-scatter + all-gather kernel over NVLink peer memory inside the step graph",
                             "nccl": "one in-place NCCL allreduce after the step graph"}[exchange],
                "per_rank_ms_per_step": [round(x, 4) for x in per_rank_ms],
                "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "kernel_ms": dms, "kernel_share_of_step": dms / (ms_total / args.steps),
                             "other_kernel_ms": {names[n]: prof[n][0] / max(prof[n][1], 1) for n in prof if n != dom},
                             "hit_pixel_fraction": hit,
                             "pixel_pass_GBps": 36.0 * S * S * B / (sum(prof[n][0] / max(prof[n][1], 1) for n in prof) * 1e-3) / 1e9,
                             "whole_step_GBps": (36.0 * S * S + 36.0 * hc.V) * B / (ms_total / args.steps * 1e-3) / 1e9,
                             "note": "ALU/latency-bound rasterisation; HBM fraction reported as the metric demands"},
                "counters": cnt, "losses": {k: round(v, 6) for k, v in losses.items()}}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(4, 1, frames=4, size=S)  # a bounded sample: 5 x 4 frames, about 10 s of host time
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        if world == 1 and not args.no_standin and args.model == "hand":
            try:  # the same eager PyTorch program on the GPU: stand-in for the reference's 1-GPU PyTorch3D path
                r = cpu_reference(5, 2, frames=4, size=S, device=f"cuda:{local}")
                line["gpu_eager_standin"] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "sample": r["sample"],
                                             "what": "oracle (reference loop body on the restated PyTorch3D ops) in eager PyTorch on this "
                                                     "GPU; NOT PyTorch3D's CUDA rasteriser, which cannot be installed offline",
                                             "ratio_value_over_standin": value / r["value"]}
            except Exception as e:  # pragma: no cover
                import traceback
                traceback.print_exc()
                line["gpu_eager_standin"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
            try:
                torch.cuda.empty_cache()
                line["dropin_level1"] = dropin_leg(sc, hc, B, S, steps=5, warmup=2)
                line["dropin_level1"]["fused_over_dropin"] = value / line["dropin_level1"]["value"]
            except Exception as e:  # pragma: no cover
                line["dropin_level1"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
