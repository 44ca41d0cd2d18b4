#!/usr/bin/env python3
"""Build every native component in-tree for sm_100a.

Outputs land in ``batch_shipyard_b200/_native/`` (git-ignored, but shipped to
the GPU box by gpurun).  Incremental: a target is rebuilt only when one of its
sources/headers is newer than the output.  ``python native/build.py [-f] [names]``.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "native")
OUT = os.path.join(ROOT, "batch_shipyard_b200", "_native")
OBJ = os.path.join(ROOT, "build", "obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "--extended-lambda",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas", "-Xptxas", "-v"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread"]


TARGETS = {
    # name: (kind, output, sources, extra flags, link flags)
    "coll": ("nvcc-shared", "libshipyard_coll.so",
             ["coll/kernels.cu", "coll/comm.cpp", "coll/gpu_mem.cpp", "coll/stub.cpp", "coll/bootstrap.cpp"],
             ["-I", os.path.join(NATIVE, "coll")], ["-lrt", "-lpthread"]),
    "preload": ("nvcc-shared", "libshipyard_preload.so",
                ["coll/preload_nccl.cpp"],
                ["-I", os.path.join(NATIVE, "coll"), "-I", "/usr/include"], ["-ldl"]),
    "mpi": ("nvcc-shared", "libshipyard_mpi.so",
            ["coll/mpi_face.cpp"],
            ["-I", os.path.join(NATIVE, "coll"), "-I", os.path.join(NATIVE, "include")], ["-ldl"]),
    "gemm": ("nvcc-shared", "libshipyard_gemm.so",
             ["gemm/gemm_tcgen05.cu"], ["-I", os.path.join(NATIVE, "coll")], ["-lcuda"]),
    "ops": ("nvcc-shared", "libshipyard_ops.so",
            ["ops/fused_ops.cu"], [], []),
    "stage": ("nvcc-shared", "libshipyard_stage.so",
              ["stage/stage.cpp"], [], ["-lpthread"]),
    "taskrun": ("cxx-exe", "shipyard-taskrun", ["runner/taskrun.cpp"], [], ["-lpthread"]),
    "gpuprobe": ("nvcc-exe", "shipyard-gpuprobe", ["probe/gpuprobe.cpp"], [], ["-ldl"]),
    "mpibench": ("nvcc-exe", "shipyard-mpibench", ["bench/mpibench.cpp"],
                 ["-I", os.path.join(NATIVE, "include"), "-I", os.path.join(NATIVE, "coll")], ["-ldl", "-lshipyard_coll"]),
    "diskbench": ("cxx-exe", "shipyard-diskbench", ["bench/diskbench.cpp"], [], ["-lpthread"]),
}
# link-time dependencies between our own libraries
DEPS = {"preload": ["coll"], "mpi": ["coll"], "mpibench": ["mpi"]}


def _newer(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers() -> list[str]:
    hs = []
    for d, _, fs in os.walk(NATIVE):
        hs += [os.path.join(d, f) for f in fs if f.endswith((".h", ".cuh", ".hpp", ".inc"))]
    return hs


def _run(cmd: list[str], log: str | None = None) -> None:
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + p.stdout + "\n")
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    if os.environ.get("SHIPYARD_BUILD_VERBOSE"):
        print(f"  [{time.time() - t0:5.1f}s] {' '.join(cmd[-3:])}")


def build_target(name: str, force: bool = False) -> str:
    kind, outname, srcs, flags, ldflags = TARGETS[name]
    srcs_abs = [os.path.join(NATIVE, s) for s in srcs]
    missing = [s for s in srcs_abs if not os.path.exists(s)]
    if missing:
        return f"{name}: skipped (missing {', '.join(os.path.relpath(m, ROOT) for m in missing)})"
    out = os.path.join(OUT, outname)
    hdrs = _headers()
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    dep_libs = [os.path.join(OUT, TARGETS[d][1]) for d in DEPS.get(name, [])]
    if not force and not _newer(out, srcs_abs + hdrs + dep_libs + [os.path.abspath(__file__)]):
        return f"{name}: up to date"
    objs = []
    jobs = []
    for s in srcs_abs:
        o = os.path.join(OBJ, name + "_" + os.path.basename(s) + ".o")
        objs.append(o)
        if not force and not _newer(o, [s] + hdrs + [os.path.abspath(__file__)]):
            continue
        if kind.startswith("nvcc"):
            cmd = [NVCC] + GENCODE + NVCC_FLAGS + flags + ["-c", s, "-o", o]
            if not s.endswith(".cu"):
                cmd = [NVCC] + GENCODE + ["-O3", "-std=c++17", "-x", "cu", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"] \
                    + flags + ["-c", s, "-o", o]
        else:
            cmd = [CXX] + CXX_FLAGS + flags + ["-c", s, "-o", o]
        jobs.append((cmd, o + ".log"))
    with cf.ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs)))) as ex:
        list(ex.map(lambda j: _run(*j), jobs))
    rpath = ["-Xlinker", "-rpath,$ORIGIN"] if kind.startswith("nvcc") else ["-Wl,-rpath,$ORIGIN"]
    deplink = []
    for d in DEPS.get(name, []):
        deplink += ["-L", OUT, "-l" + TARGETS[d][1][3:-3]]
    tmp = out + ".tmp%d" % os.getpid()          # link to a temp name, then rename: readers never see a half-written artefact
    if kind == "nvcc-shared":
        _run([NVCC] + GENCODE + ["-shared", "-o", tmp] + objs + deplink + rpath + ldflags)
    elif kind == "nvcc-exe":
        _run([NVCC] + GENCODE + ["-o", tmp] + objs + deplink + rpath + ldflags)
    else:
        _run([CXX, "-o", tmp] + objs + deplink + rpath + ldflags)
    os.replace(tmp, out)
    return f"{name}: built {os.path.relpath(out, ROOT)}"


def build_all(names: list[str] | None = None, force: bool = False, quiet: bool = False) -> None:
    order = ["coll", "preload", "mpi", "gemm", "ops", "stage", "taskrun", "gpuprobe", "mpibench", "diskbench"]
    names = names or order
    for n in order:
        if n in names:
            msg = build_target(n, force)
            if not quiet:
                print(msg)


def artifact(name: str) -> str:
    """Path of a built artefact (raises if it is not built)."""
    p = os.path.join(OUT, TARGETS[name][1])
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not built; run `python native/build.py {name}`")
    return p


def build_sanitizers() -> dict:
    """Sanitizer builds of the host-side native code: the multi-threaded staging library under
    ThreadSanitizer with a concurrent driver, the task runner under Address+UndefinedBehaviour sanitizers.
    Outputs go to build/san/ (never shipped).  Returns {name: path}."""
    san = os.path.join(ROOT, "build", "san")
    os.makedirs(san, exist_ok=True)
    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    out = {"stage_tsan": os.path.join(san, "stage_tsan"), "taskrun_asan": os.path.join(san, "shipyard-taskrun-asan")}
    # the sanitizer runtimes ship with the distribution compiler, not necessarily with $CXX
    cxx = os.environ.get("CXX_SAN") or next((c for c in ("/usr/bin/g++", "/usr/bin/clang++") if os.path.exists(c)), CXX)
    _run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-I", cuda_inc,
          os.path.join(NATIVE, "tests", "tsan_stage.cpp"), os.path.join(NATIVE, "stage", "stage.cpp"),
          "-o", out["stage_tsan"], "-L", cuda_lib, "-lcudart", "-Wl,-rpath," + cuda_lib], os.path.join(san, "stage_tsan.log"))
    _run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-pthread",
          os.path.join(NATIVE, "runner", "taskrun.cpp"), "-o", out["taskrun_asan"]], os.path.join(san, "taskrun_asan.log"))
    return out


if __name__ == "__main__":
    if "--sanitizers" in sys.argv:
        for k, v in build_sanitizers().items():
            print(f"{k}: built {os.path.relpath(v, ROOT)}")
        sys.exit(0)
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    if "--clean" in sys.argv:
        shutil.rmtree(OBJ, ignore_errors=True)
    build_all(args or None, force="-f" in sys.argv)
