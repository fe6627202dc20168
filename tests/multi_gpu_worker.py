"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with torch.distributed.run): the frame-sharded step on
R ranks must equal the single-GPU step on the same global batch — shared-parameter gradients and losses after the exchange,
per-frame gradients row by row, and the parameters after captured-graph replays with Adam — for both exchange paths
(NCCL allreduce after the graph; the peer-memory exchange kernel inside the graph)."""
import json
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from util import load_golden, golden_case, rel_l2  # noqa: E402

PER_FRAME = ("pose", "rot", "trans", "cam", "light_positions")


def fitter(hc, params, batch, cfg, rows, world, mode):
    from harp_b200.step import HarpFitter
    T = params["pose"].shape[0]
    p = dict(params)
    p["light_positions"] = params["light_position"][None].repeat(T, 1)
    p = {k: (v[rows] if k in PER_FRAME else v) for k, v in p.items()}
    frames = {k: batch[k][rows] for k in ("y_true", "y_sil", "y_sil_eroded", "init_joints")}
    return HarpFitter(hc, p, frames, batch["ref_verts"], batch["uv_mask"], cfg["img_size"], cfg["focal_length"], world_size=world,
                      exchange=mode)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from harp_b200 import synth
    from harp_b200.ops import HarpContext
    g = load_golden("c2_full_128_sub")
    model, topo, params, batch, cfg = golden_case(g)
    T = params["pose"].shape[0]
    assert T % world == 0
    per = T // world
    rows = slice(rank * per, (rank + 1) * per)
    oa, on = batch["off_albedo"].to(dev).int().contiguous(), batch["off_normal"].to(dev).int().contiguous()
    tmpl = synth.load_template("hand")
    hc = HarpContext(synth.make_mano_pack(seed=0, template=tmpl)["th_faces"], subdivide=True, faces_uvs=tmpl["faces_uvs"],
                     verts_uvs=tmpl["verts_uvs"], mano=synth.make_mano_pack(seed=0, template=tmpl), device=dev)
    out = {}
    # single-GPU reference on the whole batch (every rank computes it: the comparison needs no further communication)
    ref = fitter(hc, params, batch, cfg, slice(0, T), 1, "nccl")
    fid_all = torch.arange(T, device=dev)
    ref.step(fid_all, oa, on, stage="both", update=False)
    ref_g = {k: v.clone() for k, v in {**ref.coarse.g, **ref.app.g}.items()}
    ref_l = ref.losses.clone()
    ref3 = fitter(hc, params, batch, cfg, slice(0, T), 1, "nccl")
    for _ in range(3):
        ref3.step(fid_all, oa, on, stage="both", update=True)
    fails = []
    for mode in ("nccl", "peer"):
        ft = fitter(hc, params, batch, cfg, rows, world, mode)
        fid = torch.arange(per, device=dev)
        ft.step(fid, oa, on, stage="both", update=False)
        torch.cuda.synchronize()
        mine = {**ft.coarse.g, **ft.app.g}
        errs = {}
        for k in ("verts_disps", "shape", "texture", "normal_map", "amb_ratio"):
            errs[k] = rel_l2(mine[k].cpu(), ref_g[k].cpu())
        errs["light0"] = rel_l2(mine["light_positions"][0].cpu(), ref_g["light_positions"][0].cpu())
        for k in ("pose", "cam"):
            errs[k + "_rows"] = rel_l2(mine[k].cpu(), ref_g[k][rows].cpu())
        errs["losses"] = rel_l2(ft.losses.cpu(), ref_l.cpu())
        # captured graph + Adam: 1 warm-up step + 2 replays = 3 updates
        ft3 = fitter(hc, params, batch, cfg, rows, world, mode)
        replay = ft3.capture(fid, oa, on, stage="both", warmup=1)
        for _ in range(2):
            replay()
        torch.cuda.synchronize()
        errs["params_shared_maxabs"] = float((ft3.coarse.p["verts_disps"] - ref3.coarse.p["verts_disps"]).abs().max())
        errs["texture_maxabs"] = float((ft3.app.p["texture"] - ref3.app.p["texture"]).abs().max())
        errs["pose_rows_maxabs"] = float((ft3.coarse.p["pose"] - ref3.coarse.p["pose"][rows]).abs().max())
        out[mode] = errs
        for k, v in errs.items():
            lim = 2e-3 if k.endswith("maxabs") else (2e-3 if k in ("texture", "amb_ratio") else 2e-4)
            if not v <= lim:
                fails.append((mode, k, v))
        ft.close(); ft3.close()
        dist.barrier()
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "errors": out, "fails": fails})
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_parity_n{world}.json"), "w") as f:
            json.dump(gathered, f, indent=1)
        print(json.dumps(gathered))
    dist.destroy_process_group()
    if any(x["fails"] for x in gathered):
        sys.exit(1)


if __name__ == "__main__":
    main()
