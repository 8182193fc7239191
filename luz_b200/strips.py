"""Image-space partition of one frame over the GPUs of a box:
host-side mirror of the row arithmetic in csrc/api.cu (luzrt_resize / luzrt_set_gbuffer / luzrt_light_pass /
luzrt_gather), used by bench.py and the multi-process tests.

The frame is cut into bands of band_rows(H, world) rows; band b belongs to rank b % world (round-robin, so sky
and geometry are spread over the ranks).  A rank shades its bands plus one extra row above and below each
(wrapping at the image border like the reference's REPEAT sampler, VulkanWrapper.cpp:2433-2437) because
taa.comp's 3x3 taps read them (taa.comp:33-41, :93-103).  The RGBA32F light images are stored with a rank's rows
contiguous (storage_row), so the all-gather of the resolved frame is one in-place collective."""
import os


def band_rows(height, world):
    if world < 1 or world & (world - 1):
        raise ValueError("the number of ranks must be a power of two, got %d" % world)
    if height % world:
        raise ValueError("height %d is not divisible by %d ranks" % (height, world))
    rpr = height // world
    if world == 1:
        return height
    if rpr < 2:
        raise ValueError("a rank needs at least two rows")
    best = rpr
    min_rows = max(int(os.environ.get("LUZRT_BAND_ROWS_MIN", "48")), 2)  # same knob as csrc/api.cu
    for k in range(1, rpr + 1):  # smallest band >= min_rows rows that divides a rank's share
        if rpr % k == 0 and rpr // k >= min_rows:
            best = rpr // k
    return best


def owned_bands(rank, world, height):
    """[(first_row, end_row)] of the bands rank owns, top to bottom."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    hb = band_rows(height, world)
    return [(b * hb, (b + 1) * hb) for b in range(rank, height // hb, world)]


def owned_rows(rank, world, height):
    return [y for lo, hi in owned_bands(rank, world, height) for y in range(lo, hi)]


def shaded_rows(rank, world, height):
    """Rows the light pass evaluates on this rank, in kernel order (may wrap): each own band + 1 halo row each side."""
    if world == 1:
        return list(range(height))
    return [y % height for lo, hi in owned_bands(rank, world, height) for y in range(lo - 1, hi + 1)]


def upload_segments(rank, world, height):
    """Contiguous [lo, hi) row ranges of the full-frame G-buffer this rank needs (what luzrt_set_gbuffer copies)."""
    if world == 1:
        return [(0, height)]
    segs = []
    for lo, hi in owned_bands(rank, world, height):
        if lo - 1 < 0:
            segs.append((height - 1, height))
        if hi + 1 > height:
            segs.append((0, 1))
        segs.append((max(lo - 1, 0), min(hi + 1, height)))
    return segs


def h2d_bytes(rank, world, width, height, scene_block_bytes=31200):
    """Bytes one end-to-end step uploads on this rank: 32 B/px of G-buffer over its segments + the SceneBlock."""
    rows = sum(hi - lo for lo, hi in upload_segments(rank, world, height))
    return rows * width * 32 + scene_block_bytes


def storage_row(y, world, height):
    """Row of image row y inside the band-permuted light images (rank-major, then band, then row)."""
    hb = band_rows(height, world)
    b = y // hb
    return (b % world) * (height // world) + (b // world) * hb + y % hb


def gather_layout(world, width, height):
    """(offset, count) in floats of every rank's rows inside the gathered (band-permuted) RGBA32F frame."""
    n = (height // world) * width * 4
    return [(r * n, n) for r in range(world)]
