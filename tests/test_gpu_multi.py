"""GPU tests of the multi-GPU plumbing: the content hash of the acceleration structures (determinism across contexts and
GPUs) and the single-process multi-device entry points (luzrt_create_multi / luzrt_gather_multi /
luzrt_comm_check_bvh_multi; SURVEY section 8b: Luz is one process, one thread).  The two-GPU case needs two devices and is
skipped on a one-GPU box (the driver's test box); it is run with `gpurun --gpus 2`."""
import numpy as np
import pytest

import scene_util as S
from luz_b200 import rt as R
from luz_b200 import strips

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _frame(rt, sc, w, h, bn):
    rt.resize(w, h)
    rt.set_blue_noise(bn)
    S.make_rt_scene(rt, sc)
    rt.set_scene(sc["scene"])
    rt.set_debug(0)
    rt.gbuffer_pass(sc["models"], len(sc["instances"]))
    rt.light_pass(3)
    rt.taa_pass(True)


def test_bvh_hash_is_a_function_of_the_scene(rt_factory):
    w, h = 128, 96
    sc = S.synthetic_scene(w, h, grid=4, n_lights=2)
    hashes = []
    for _ in range(3):
        rt = rt_factory()
        rt.resize(w, h)
        S.make_rt_scene(rt, sc)
        hashes.append(rt.bvh_hash())
        assert rt.bvh_hash() == hashes[-1]
        rt.close()
    assert hashes[0] == hashes[1] == hashes[2] and hashes[0] != 0
    moved = dict(sc, instances=[(m, np.asarray(mat, np.float32) + (np.arange(16) == 12) * np.float32(1e-3 * (i == 3)), ci)
                                for i, (m, mat, ci) in enumerate(sc["instances"])])
    rt = rt_factory()
    rt.resize(w, h)
    S.make_rt_scene(rt, moved)
    assert rt.bvh_hash() != hashes[0]  # one instance moved by a millimetre
    rt.close()


def test_create_multi_with_one_device_is_an_ordinary_ctx():
    (rt,) = R.create_multi([0])
    w, h = 96, 64
    sc = S.synthetic_scene(w, h, grid=2, n_lights=1)
    _frame(rt, sc, w, h, S.blue_noise())
    R.gather_multi([rt])
    R.comm_check_bvh_multi([rt])
    assert np.isfinite(rt.read(R.IMG_LIGHT)).all()
    rt.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_devices_in_one_process_assemble_the_single_gpu_frame(rt_factory):
    w, h = 256, 384
    sc = S.synthetic_scene(w, h, grid=4, n_lights=3, light_samples=1, ao_samples=6)
    bn = S.blue_noise()
    ref = rt_factory()
    _frame(ref, sc, w, h, bn)
    full = ref.read(R.IMG_LIGHT)
    ctxs = R.create_multi([0, 1])
    for c in ctxs:
        _frame(c, sc, w, h, bn)
    R.comm_check_bvh_multi(ctxs)  # both replicas of the BVH are bitwise the same
    assert ctxs[0].bvh_hash() == ctxs[1].bvh_hash() == ref.bvh_hash()
    R.gather_multi(ctxs)
    for rank, c in enumerate(ctxs):
        got = c.read(R.IMG_LIGHT)  # after the gather every GPU holds the whole resolved frame
        assert np.array_equal(got, full), "rank %d" % rank
        own = np.array(strips.owned_rows(rank, 2, h))
        assert own.size == h // 2
    for c in ctxs:
        c.close()
