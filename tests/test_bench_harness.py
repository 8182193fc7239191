"""CPU tests of the benchmark harness: the workload loader bench.py and the benchmark-scale parity tests share, the
staleness rule of the committed ncu counters, and the CPU reference arm's JSON contract (bench.py --impl reference)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from luz_b200 import scenes, workloads  # noqa: E402


def test_workload_loads_through_the_host_mirror_and_animates(tmp_path):
    wl = workloads.Workload(None, "c2", tmp=str(tmp_path))
    wl.upload(None)
    wl.app.update_resources()
    assert len(wl.app.instances()) == 4097 and wl.app.light_count() == 4 and wl.animate == "refit"
    m0 = np.array(wl.app.instances()[5][1])
    wl.move(10)  # yaw += 0.5 deg per frame for every cube, the slab stays
    wl.app.update_resources()
    m1 = np.array(wl.app.instances()[5][1])
    assert not np.array_equal(m0, m1) and np.array_equal(np.array(wl.app.instances()[0][1]), np.array(wl.app.instances()[0][1]))
    assert np.allclose(m0[12:15], m1[12:15])  # rotation in place
    sb = wl.app.scene_block()
    assert sb.ao_num_samples == scenes.CONFIGS["c2"]["ao_samples"] and sb.num_lights == 4


def test_sample_overrides_show_in_the_configuration(tmp_path):
    wl = workloads.Workload(None, "c1", tmp=str(tmp_path), light_samples=0, ao_samples=3)
    wl.upload(None)
    wl.app.update_resources()
    sb = wl.app.scene_block()
    assert sb.ao_num_samples == 3 and sb.lights[0].num_shadow_samples == 0


def test_ncu_counters_are_dropped_when_the_kernels_changed(tmp_path, monkeypatch):
    class A:
        config, variant, width, gpus = "c3", None, 0, 1
    h = bench.kernel_source_hash()
    assert h == bench.kernel_source_hash() and len(h) == 16
    doc = {"source_hash": h, "c3": {"k_taa": {"dram_bytes": 123.0, "active_lanes": 31.9}}}
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "traffic.json").write_text(json.dumps(doc))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.ncu_traffic(A, "k_taa", h) == 123.0
    assert bench.ncu_counters(A, ("k_taa",), h)["k_taa"]["active_lanes"] == 31.9
    assert bench.ncu_traffic(A, "k_taa", "0" * 16) is None  # captured from other sources: stale
    assert bench.ncu_counters(A, ("k_taa",), "0" * 16) is None
    A.gpus = 8
    assert bench.ncu_traffic(A, "k_taa", h) is None  # a 1-GPU capture says nothing about a rank's share


def test_reference_arm_prints_the_contract_line_with_every_core():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports: the arm must not inherit it
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--ref-rows", "12",
                        "--steps", "1", "--warmup", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("c1: 1280x720")
    # under torchrun only rank 0 runs it
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert q.returncode == 0 and not [l for l in q.stdout.splitlines() if l.startswith("{")]
