"""Image-space partition of one frame over the GPUs of a box: host-side mirror of the
row arithmetic in csrc/api.cu (luzrt_resize / luzrt_set_gbuffer / luzrt_light_pass / luzrt_gather), used by
bench.py and the multi-process tests.  Rank r of `world` owns rows [r*H/world, (r+1)*H/world); it shades one
extra row above and below (wrapping at the image border like the reference's REPEAT sampler,
VulkanWrapper.cpp:2433-2437) because taa.comp's 3x3 taps read them (taa.comp:33-41, :93-103)."""


def owned_rows(rank, world, height):
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    if height % world:
        raise ValueError("height %d is not divisible by %d ranks" % (height, world))
    n = height // world
    return rank * n, (rank + 1) * n


def shaded_rows(rank, world, height):
    """Rows the light pass evaluates on this rank, in kernel order (may wrap): own strip + 1 halo row each side."""
    y0, y1 = owned_rows(rank, world, height)
    if world == 1:
        return list(range(height))
    return [(y0 - 1 + i) % height for i in range(y1 - y0 + 2)]


def upload_segments(rank, world, height):
    """Contiguous [lo, hi) row ranges of the full-frame G-buffer this rank needs (what luzrt_set_gbuffer copies)."""
    y0, y1 = owned_rows(rank, world, height)
    if world == 1 or (y1 - y0) + 2 >= height:
        return [(0, height)]
    segs = [(max(y0 - 1, 0), min(y1 + 1, height))]
    if y0 == 0:
        segs.append((height - 1, height))
    if y1 == height:
        segs.append((0, 1))
    return segs


def h2d_bytes(rank, world, width, height, scene_block_bytes=31200):
    """Bytes one end-to-end step uploads on this rank: 32 B/px of G-buffer over its segments + the SceneBlock."""
    rows = sum(hi - lo for lo, hi in upload_segments(rank, world, height))
    return rows * width * 32 + scene_block_bytes


def gather_layout(world, width, height):
    """(offset, count) in floats of every rank's strip inside the gathered RGBA32F frame."""
    n = (height // world) * width * 4
    return [(r * n, n) for r in range(world)]
