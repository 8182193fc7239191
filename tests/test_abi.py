"""CPU-side checks of the drop-in boundary: libluzrt.so / libluzhost.so load, export every symbol that
include/luzrt.h declares, refuse to run without a GPU (no fallback), and never route through oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

from luz_b200 import host, rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"LUZRT_API\s+[\w\s\*]+?\b(luzrt_\w+)\s*\(", src)))


def test_header_symbols_are_all_exported():
    names = declared_symbols("luzrt.h")
    assert len(names) >= 30
    lib = C.CDLL(rt.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes harness binds exactly the header's entry points
    assert sorted(rt.EXPORTS) == names


def test_library_exports_only_the_c_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", rt.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in syms if not s.startswith(("_init", "_fini", "__"))]
    assert ours and all(s.startswith("luzrt_") for s in ours), [s for s in ours if not s.startswith("luzrt_")][:10]


def test_product_libraries_do_not_link_the_oracle():
    for p in (rt.LIB_PATH, host.LIB_PATH):
        out = subprocess.run(["ldd", p], stdout=subprocess.PIPE, text=True).stdout
        assert "luz_oracle" not in out
        with open(p, "rb") as f:
            assert b"libluz_oracle" not in f.read()
    for d in ("luz_b200",):
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for fn in files:
                # build.py only *compiles* the checker (make -C oracle); building it is not using it
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and fn != "build.py":
                    with open(os.path.join(base, fn), errors="ignore") as f:
                        src = f.read()
                    assert "oracle_api" not in src and "luz_oracle" not in src, fn


def test_create_fails_loudly_without_a_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    lib = rt.load_library()
    ctx = C.c_void_p()
    assert lib.luzrt_create(0, 0, 1, C.byref(ctx)) == -6  # LUZRT_E_NODEVICE
    assert not ctx.value
    with pytest.raises(rt.LuzError):
        rt.LuzRT(device=0)


def test_argument_validation_needs_no_device():
    lib = rt.load_library()
    assert lib.luzrt_create(0, 0, 1, None) == -1
    ctx = C.c_void_p()
    assert lib.luzrt_create(0, 3, 2, C.byref(ctx)) == -1  # rank outside world
    assert lib.luzrt_resize(None, 16, 16) == -1
    assert lib.luzrt_light_pass(None, 0) == -1
    assert lib.luzrt_launch_count(None) == 0
    assert b"sm_100a" in lib.luzrt_version()


def test_wire_header_static_layout():
    """include/luz_wire.h carries static_asserts for the sizes measured on the compiled reference; compile it as
    C and C++ to make sure they hold with this toolchain."""
    for comp, std in (("gcc", "-std=c11"), ("g++", "-std=c++17")):
        src = '#include "luz_wire.h"\n#include "luzrt.h"\nint main(void){return sizeof(luzw_scene_block)==31200?0:1;}\n'
        ext = ".c" if comp == "gcc" else ".cpp"
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "t" + ext)
            with open(p, "w") as f:
                f.write(src)
            exe = os.path.join(d, "t")
            subprocess.check_call([comp, std, "-I", os.path.join(ROOT, "include"), p, "-o", exe])
            assert subprocess.call([exe]) == 0
