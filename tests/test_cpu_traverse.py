"""The traversal logic of the CUDA kernels, checked WITHOUT a GPU: luz_b200/csrc/traverse.cuh is compiled for the host
(tests/cpu_traverse/harness.cpp shims the device intrinsics; the header's three inline-PTX sequences have host branches)
and run over wide BVHs built on the CPU in the kernels' node format, against an exhaustive double-precision
ray/triangle test.  Covers trace_ray any-hit and closest-hit, the candidate-list path (collect_instances +
filter_candidates), occluder hints (then_root) and the Pluecker triangle test, on rotated, anisotropically scaled and
mirrored instances.  This is test infrastructure; the product never traverses on the CPU."""
import json
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers (vector types only)")
    exe = str(tmp_path_factory.mktemp("cpu_traverse") / "harness")
    subprocess.check_call([gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-attributes", "-I", CUDA_INC,
                           "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(HERE, "cpu_traverse", "harness.cpp")])
    return exe


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_host_compiled_traversal_matches_exhaustive(harness, seed):
    r = json.loads(subprocess.run([harness, str(seed)], stdout=subprocess.PIPE, text=True, check=True).stdout)
    assert r["rays"] == 6000 and 0.1 * r["rays"] < r["hits"] < 0.9 * r["rays"], r  # a real mix of hits and misses
    assert r["nodes_visited"] > 3 * r["rays"] and r["tris_tested"] > 0 and r["tlas_levels"] >= 2
    assert r["agree"] >= 0.999 * r["rays"], r              # the north star's visibility gate
    assert r["clear_agree"] == r["clear_rays"], r          # every ray that does not graze an edge agrees
    assert r["hint_same"] == r["rays"], r                  # an occluder hint never changes the answer
    assert r["cand_rays"] > 1000 and r["cand_same"] == r["cand_rays"], r  # candidate lists == root descent
    assert r["closest_ok"] == r["closest_n"], r            # closest hit: same hit / miss and t within 1e-4


def test_host_compiled_traversal_deep_tlas(harness):
    """600 instances in a denser box: a four-level TLAS, long stacks, most rays occluded."""
    r = json.loads(subprocess.run([harness, "7", "600", "2500", "9"], stdout=subprocess.PIPE, text=True, check=True).stdout)
    assert r["rays"] == 2500 and r["tlas_levels"] >= 4 and r["hits"] > 0.3 * r["rays"], r
    assert r["agree"] >= 0.999 * r["rays"] and r["clear_agree"] == r["clear_rays"], r
    assert r["hint_same"] == r["rays"] and r["cand_same"] == r["cand_rays"] and r["closest_ok"] == r["closest_n"], r
