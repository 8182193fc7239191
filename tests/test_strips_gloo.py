"""Multi-rank host logic on CPU (gloo, world_size 2): the image-strip partition (luz_b200/strips.py mirrors
csrc/api.cu), the halo rows TAA needs, strip assembly by all-gather, and bench.py's max-over-ranks reduction.
The per-strip 'compute' here is the CPU oracle standing in for the CUDA passes (test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_api as O
import scene_util as S
from luz_b200 import strips


@pytest.mark.parametrize("world,height", [(1, 720), (2, 720), (4, 1080), (8, 2160), (8, 4320), (2, 4)])
def test_partition_covers_image_once(world, height):
    seen = np.zeros(height, np.int32)
    for r in range(world):
        y0, y1 = strips.owned_rows(r, world, height)
        seen[y0:y1] += 1
        sh = strips.shaded_rows(r, world, height)
        if world > 1:
            assert len(sh) == (y1 - y0) + 2 and sh[0] == (y0 - 1) % height and sh[-1] == y1 % height
        # every shaded row is covered by an uploaded G-buffer segment
        up = np.zeros(height, bool)
        for lo, hi in strips.upload_segments(r, world, height):
            assert 0 <= lo < hi <= height
            up[lo:hi] = True
        assert up[np.array(sh)].all()
    assert (seen == 1).all()
    lay = strips.gather_layout(world, 16, height)
    assert lay[0][0] == 0 and lay[-1][0] + lay[-1][1] == 16 * height * 4


def test_partition_rejects_bad_arguments():
    with pytest.raises(ValueError):
        strips.owned_rows(2, 2, 720)
    with pytest.raises(ValueError):
        strips.owned_rows(0, 7, 720)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O.lib().orc_set_threads(2)
        sc = S.synthetic_scene(w, h, grid=3, n_lights=2, light_samples=1, ao_samples=2)
        world_geo = O.World(sc["meshes"], sc["instances"])
        gb = O.gbuffer_pass(sc["scene"], world_geo, sc["models"], len(sc["instances"]), [], w, h, exhaustive=False)
        bn = S.blue_noise()
        y0, y1 = strips.owned_rows(rank, world, h)
        # light pass over the rows this rank shades (own strip + wrapped halo rows), row range by row range
        light = np.zeros((h, w, 4), np.float32)
        rays = 0
        for y in strips.shaded_rows(rank, world, h):
            rc, out, _, _, st = O.light_pass(sc["scene"], gb, 5, bn, world_geo, exhaustive=False, rows=(y, y + 1))
            assert rc == 0
            light[y] = out[y]
            if y0 <= y < y1:
                rays += st.rays  # halo rows are recomputation, not frame rays
        # frame 0 convention: history == current light buffer; the gathered frame is the next history
        resolved = O.taa_pass(sc["scene"], light, light, gb.depth, True, rows=(y0, y1))
        mine = torch.from_numpy(np.ascontiguousarray(resolved[y0:y1]))
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        frame = torch.cat(parts, 0).numpy()
        # bench.py's reductions: time = max over ranks, rays = sum over ranks
        t = torch.tensor([10.0 + rank, float(rays)], dtype=torch.float64)
        mx, sm = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        if rank == 0:
            rc, full, _, _, st = O.light_pass(sc["scene"], gb, 5, bn, world_geo, exhaustive=False)
            ref = O.taa_pass(sc["scene"], full, full, gb.depth, True)
            q.put(dict(equal=bool(np.array_equal(frame, ref)), max_ms=float(mx[0]), rays=float(sm[1]),
                       ref_rays=float(st.rays), nonzero=float(np.abs(ref).sum())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_strips_assemble_bitwise_equal_frame():
    world, w, h = 2, 64, 36
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["equal"], "strips gathered from 2 ranks differ from the 1-rank frame"
    assert res["max_ms"] == 11.0 and res["rays"] == res["ref_rays"] and res["rays"] > 0 and res["nonzero"] > 0
