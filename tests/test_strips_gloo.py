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


@pytest.mark.parametrize("world,height", [(1, 720), (2, 720), (4, 1080), (8, 2160), (8, 4320), (2, 4), (8, 720)])
def test_partition_covers_image_once(world, height):
    seen = np.zeros(height, np.int32)
    hb = strips.band_rows(height, world)
    assert (height // world) % hb == 0 and (world == 1 or hb >= 48 or hb == height // world)
    for r in range(world):
        own = strips.owned_rows(r, world, height)
        seen[np.array(own)] += 1
        sh = strips.shaded_rows(r, world, height)
        if world > 1:
            assert len(sh) == len(own) + 2 * len(strips.owned_bands(r, world, height))
            assert set(own) <= set(sh)
            for lo, hi in strips.owned_bands(r, world, height):
                assert (lo - 1) % height in sh and hi % height in sh
        # every shaded row is covered by an uploaded G-buffer segment
        up = np.zeros(height, bool)
        for lo, hi in strips.upload_segments(r, world, height):
            assert 0 <= lo < hi <= height
            up[lo:hi] = True
        assert up[np.array(sh)].all()
        # a rank's rows are contiguous in storage order, in band order
        st = [strips.storage_row(y, world, height) for y in own]
        assert st == list(range(r * (height // world), (r + 1) * (height // world)))
    assert (seen == 1).all()
    assert sorted(strips.storage_row(y, world, height) for y in range(height)) == list(range(height))
    lay = strips.gather_layout(world, 16, height)
    assert lay[0][0] == 0 and lay[-1][0] + lay[-1][1] == 16 * height * 4


def test_bands_balance_a_sky_over_ground_frame():
    """Contiguous strips give one rank all the sky; round-robin bands keep every rank within a few percent."""
    h, world = 2160, 8
    cost = np.where(np.arange(h) < 900, 0.05, 1.0)  # top 900 rows: background, no rays
    per_rank = [cost[np.array(strips.owned_rows(r, world, h))].sum() for r in range(world)]
    assert max(per_rank) / (sum(per_rank) / world) < 1.15


def test_partition_rejects_bad_arguments():
    with pytest.raises(ValueError):
        strips.owned_bands(2, 2, 720)
    with pytest.raises(ValueError):
        strips.band_rows(720, 7)
    with pytest.raises(ValueError):
        strips.band_rows(6, 4)


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
        own = strips.owned_rows(rank, world, h)
        # light pass over the rows this rank shades (own bands + wrapped halo rows), row by row
        light = np.zeros((h, w, 4), np.float32)
        rays = 0
        for y in strips.shaded_rows(rank, world, h):
            rc, out, _, _, st = O.light_pass(sc["scene"], gb, 5, bn, world_geo, exhaustive=False, rows=(y, y + 1))
            assert rc == 0
            light[y] = out[y]
            if y in own:
                rays += st.rays  # halo rows are recomputation, not frame rays
        # frame 0 convention: history == current light buffer; the gathered frame is the next history
        resolved = np.zeros((h, w, 4), np.float32)
        for lo, hi in strips.owned_bands(rank, world, h):
            resolved[lo:hi] = O.taa_pass(sc["scene"], light, light, gb.depth, True, rows=(lo, hi))[lo:hi]
        mine = torch.from_numpy(np.ascontiguousarray(resolved[np.array(own)]))  # band order == storage order
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        banded = torch.cat(parts, 0).numpy()  # the band-permuted frame every rank now holds
        frame = banded[np.array([strips.storage_row(y, world, h) for y in range(h)])]
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
    world, w, h = 2, 64, 192  # 2 ranks x 2 bands of 48 rows
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
