"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU path — frame sharding and the single allreduce of the
shared-parameter gradients (per-frame rows must stay rank-local)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, V):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from harp_b200.step import _Flat, allreduce_shared, shard_frames, LOSS_NAMES
        a, b = shard_frames(T, rank, world)
        n = b - a
        coarse = _Flat([("pose", torch.zeros(n, 45)), ("cam", torch.zeros(n, 3)), ("verts_disps", torch.zeros(V, 1)),
                        ("shape", torch.zeros(10))], "cpu")
        app = _Flat([("light_positions", torch.zeros(n, 3)), ("amb_ratio", torch.zeros(1)), ("texture", torch.zeros(1, 8, 8, 3)),
                     ("normal_map", torch.zeros(1, 8, 8, 3))], "cpu")
        coarse.grad.fill_(float(rank + 1))
        app.grad.fill_(float(10 * (rank + 1)))
        losses = torch.full((len(LOSS_NAMES),), float(rank + 1))
        allreduce_shared(coarse, app, losses, world)
        tot = sum(range(1, world + 1))
        assert torch.all(coarse.g["pose"] == rank + 1) and torch.all(coarse.g["cam"] == rank + 1)          # rank-local rows
        assert torch.all(coarse.g["verts_disps"] == tot) and torch.all(coarse.g["shape"] == tot)             # shared
        assert torch.all(app.g["light_positions"][0] == 10 * tot)                                           # shared light = row 0
        if n > 1:
            assert torch.all(app.g["light_positions"][1:] == 10 * (rank + 1))
        assert torch.all(app.g["amb_ratio"] == 10 * tot) and torch.all(app.g["texture"] == 10 * tot)
        assert torch.all(app.g["normal_map"] == 10 * tot)
        assert torch.allclose(losses, torch.full_like(losses, tot / world))
        # the copy-free variant HarpFitter uses: [coarse grads | losses | app grads with the rank-local light rows last] in
        # one buffer, one in-place allreduce of the shared span, whose length must not depend on the shard size n
        from harp_b200.step import allreduce_shared_inplace
        nc, nl, na = n * 48 + V + 10, len(LOSS_NAMES), 1 + 2 * 8 * 8 * 3 + n * 3
        gbuf = torch.zeros(nc + nl + na)
        c2 = _Flat([("pose", torch.zeros(n, 45)), ("cam", torch.zeros(n, 3)), ("verts_disps", torch.zeros(V, 1)),
                    ("shape", torch.zeros(10))], "cpu", grad_storage=gbuf[:nc])
        a2 = _Flat([("amb_ratio", torch.zeros(1)), ("texture", torch.zeros(1, 8, 8, 3)), ("normal_map", torch.zeros(1, 8, 8, 3)),
                    ("light_positions", torch.zeros(n, 3))], "cpu", grad_storage=gbuf[nc + nl:])
        c2.grad.fill_(float(rank + 1))
        a2.grad.fill_(float(10 * (rank + 1)))
        gbuf[nc:nc + nl] = float(rank + 1)
        start = c2.offsets[c2.names.index("verts_disps")]
        stop = nc + nl + a2.offsets[a2.names.index("light_positions")] + 3
        assert stop - start == V + 10 + nl + 1 + 2 * 8 * 8 * 3 + 3
        allreduce_shared_inplace(gbuf, start, stop, slice(nc, nc + nl), world)
        assert torch.all(c2.g["pose"] == rank + 1) and torch.all(c2.g["cam"] == rank + 1)
        assert torch.all(c2.g["verts_disps"] == tot) and torch.all(c2.g["shape"] == tot)
        assert torch.all(a2.g["light_positions"][0] == 10 * tot)
        if n > 1:
            assert torch.all(a2.g["light_positions"][1:] == 10 * (rank + 1))
        assert torch.all(a2.g["amb_ratio"] == 10 * tot) and torch.all(a2.g["texture"] == 10 * tot) and torch.all(a2.g["normal_map"] == 10 * tot)
        assert torch.allclose(gbuf[nc:nc + nl], torch.full((nl,), tot / world))
    finally:
        dist.destroy_process_group()


def test_allreduce_shared_gloo_world2():
    mp.spawn(_worker, args=(2, _free_port(), 7, 11), nprocs=2, join=True)


def test_shard_frames_partitions_exactly():
    from harp_b200.step import shard_frames
    for T in (1, 7, 64, 256, 257):
        for R in (1, 2, 4, 8):
            spans = [shard_frames(T, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(spans[i][1] == spans[i + 1][0] for i in range(R - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
