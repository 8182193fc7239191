"""In-tree build of the native libraries.

  luz_b200/libluzrt.so    CUDA kernels + C ABI (include/luzrt.h), sm_100a only
  luz_b200/libluzhost.so  C++ host mirror of Luz's GPUScene / DeferredRenderer / scene format
  oracle/libluz_oracle.so CPU oracle (test infrastructure; `make -C oracle`)

nvcc cross-compiles without a GPU.  Objects are rebuilt only when a source or header is newer.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "luz_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
BUILD = os.path.join(ROOT, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
]
CU_SOURCES = ["api.cu", "bvh_build.cu", "light_pass.cu", "taa.cu", "gbuffer.cu", "volumetric.cu", "shadow_map.cu", "relaxed.cu"]
# The shading / resolve kernels restate GLSL that the oracle evaluates without FMA contraction; they are
# built with -fmad=false so that implicit contraction cannot change results (ill-conditioned BRDF terms
# amplify it), and use explicit fmaf() only where rounding is not part of parity (box tests).
NO_FMAD = {"light_pass.cu", "taa.cu", "gbuffer.cu", "volumetric.cu", "shadow_map.cu"}
HOST_SOURCES = ["json.cpp", "scene.cpp", "gpu_scene.cpp", "capi.cpp", "import.cpp", "png.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers(d):
    out = []
    for base in (d, os.path.join(ROOT, "include")):
        for f in os.listdir(base):
            if f.endswith((".h", ".hpp", ".cuh")):
                out.append(os.path.join(base, f))
    return out


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return p.stdout


def build_luzrt(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdrs = _headers(CSRC)
    objs, jobs = [], []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append((NVCC_FLAGS_CMD(s, o, ["-fmad=false"] if src in NO_FMAD else []),
                         os.path.join(BUILD, src + ".log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(lambda j: _run(j[0], j[1]), jobs):
                if verbose:
                    print(out)
    lib = os.path.join(PKG, "libluzrt.so")
    if force or jobs or _newer(lib, objs):
        _run([NVCC, "-shared", "-o", lib] + objs + ["-ccbin", HOST_CXX, "-ldl", "-lcudart_static", "-lrt", "-lpthread"])
    return lib


def NVCC_FLAGS_CMD(src, obj, extra=()):
    # LUZ_EXTRA_NVCC: extra flags (e.g. -DLUZ_FAST_RAYGEN=0) for the A/B variants under build/variants
    return [NVCC] + NVCC_FLAGS + list(extra) + os.environ.get("LUZ_EXTRA_NVCC", "").split() + ["-c", src, "-o", obj]


def build_luzhost(force=False):
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES if os.path.exists(os.path.join(HOST, s))]
    if not srcs:
        return None
    lib = os.path.join(PKG, "libluzhost.so")
    if force or _newer(lib, srcs + _headers(HOST)):
        _run([HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unused-function",
              "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-o", lib] + srcs +
             ["-L", PKG, "-lluzrt", "-Wl,-rpath,$ORIGIN", "-ldl"])
    return lib


def build_oracle(with_ref=True):
    """The checker, not the product: libluz_oracle.so and (if the reference is mounted) oracle/_ref."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "libluz_oracle.so"])
    if with_ref and os.path.isdir("/root/reference/source"):
        _run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])


def build_all(force=False, verbose=False):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        raise RuntimeError("nvcc not found: libluzrt.so has no CPU fallback and cannot be built without it")
    build_luzrt(force, verbose)
    build_luzhost(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", [f for f in os.listdir(PKG) if f.endswith(".so")])
