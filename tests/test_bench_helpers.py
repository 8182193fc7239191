"""Host-side helpers of bench.py that run without a GPU."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("luz_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_sysfs_bus_id_from_nvidia_smi_csv():
    csv = "0, 00000000:1B:00.0\n1, 00000000:43:00.0\n7, 00000001:E4:00.0\n"
    assert bench.sysfs_bus_id(csv, 0) == "0000:1b:00.0"
    assert bench.sysfs_bus_id(csv, 7) == "0001:e4:00.0"
    assert bench.sysfs_bus_id(csv, 3) is None
    assert bench.sysfs_bus_id("garbage\n", 0) is None
    assert bench.sysfs_bus_id("0, 0000:1b:00.0\n", 0) == "0000:1b:00.0"


def test_cpulist_parsing():
    assert bench.cpulist_to_set("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert bench.cpulist_to_set("5") == {5}
    assert bench.cpulist_to_set("") == set()


def test_numa_binding_is_best_effort_and_leaves_affinity_alone_on_failure(monkeypatch):
    before = os.sched_getaffinity(0)
    note = bench.bind_to_gpu_numa_node(0)  # no nvidia-smi / no GPU here
    assert isinstance(note, str) and os.sched_getaffinity(0) == before
    monkeypatch.setenv("LUZ_BENCH_NO_NUMA", "1")
    assert bench.bind_to_gpu_numa_node(0) == "disabled"
