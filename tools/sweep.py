"""Secondary measurements next to bench.py's headline (BASELINE.json configs 0 and 4):
  C1: single 256x256 frame, 778-vert MANO (no subdivision), silhouette-only loss, fwd+bwd
  C5: forward-only raster throughput (silhouette + shadowed colour render) at 512x512 for batch 1..1024
Each line is one JSON object; run on the GPU box:  python tools/sweep.py > gpurun_out/sweep.jsonl"""
import json
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from harp_b200 import camera, scene  # noqa: E402
from harp_b200.ops import LbsMano, MeshBuild, render_forward  # noqa: E402
from harp_b200.step import HarpFitter  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, warm=3, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def c1():
    hc, _, _ = scene.make_context(subdivide=False)
    sc = scene.make_scene(hc, 1, 256)
    ft = HarpFitter(hc, sc["params"], sc["frames"], sc["ref_verts"], sc["uv_mask"], 256, sc["focal"])
    fid = torch.arange(1, device="cuda")
    run = ft.capture(fid, sc["offsets"]["albedo"], sc["offsets"]["normal"], stage="coarse", frame_range=(0, 1))
    ms = timeit(run, 5, 50)
    print(json.dumps({"config": "C1 single 256x256 frame, MANO 778 verts, silhouette+kps+mesh regs (coarse stage), fwd+bwd+Adam",
                      "ms_per_step": ms, "frames_per_s": 1e3 / ms, "note": "latency-bound: one frame cannot fill 148 SMs"}), flush=True)


def c5():
    hc, _, _ = scene.make_context()
    S = 512
    focal = 1000.0 * S / 224.0
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        try:
            sc = scene.make_scene(hc, min(B, 64), S)
            p = {k: v.cuda() for k, v in sc["params"].items()}
            rep = (B + 63) // 64
            rows = lambda t: t.repeat((rep,) + (1,) * (t.dim() - 1))[:B].contiguous()
            pose48 = rows(torch.cat((p["rot"], p["pose"]), 1))
            trans, cam = rows(p["trans"]), rows(p["cam"])
            light = p["light_positions"][0][None].expand(B, 3).contiguous()
            amb = torch.sigmoid(p["amb_ratio"])
            cR = camera.flip_xy_R(B, "cuda")

            @torch.no_grad()
            def fwd():
                verts0, _ = LbsMano.apply(hc, pose48, p["shape"], trans)
                verts, normals = MeshBuild.apply(hc, verts0, p["verts_disps"])
                cT = camera.cam_to_T(cam, focal, S)
                lR, lT = camera.light_camera(light, verts.mean(1))
                return render_forward(hc, verts, normals, cR, cT, lR, lT, light, amb, p["texture"], p["normal_map"], focal, S, 1)
            ms = timeit(fwd, 2, 10 if B < 256 else 4)
            gbps = 16.0 * S * S * B / (ms * 1e-3) / 1e9  # algorithmic: write y_pred 12 S^2 + alpha 4 S^2 per frame
            print(json.dumps({"config": f"C5 forward-only raster @512x512, batch {B}", "batch": B, "ms": ms, "frames_per_s": B * 1e3 / ms,
                              "achieved_GBps": gbps, "frac_of_hbm_peak": gbps / PEAK}), flush=True)
            del sc, p
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"config": f"C5 batch {B}", "error": str(e)[:200]}), flush=True)
            break


if __name__ == "__main__":
    c1()
    c5()
