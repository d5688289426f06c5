"""In-tree build of libbpmf_b200.so (hand-written CUDA for sm_100a, nvcc cross-compiles without a GPU).

    python -m seismic_bpmf_b200.build [--force] [--verbose]

The shared object lands in seismic_bpmf_b200/lib/ (git-ignored, travels with gpurun snapshots),
together with a ``libc.so`` symlink so that the directory can be used as ``BPMF/lib`` by the
reference's own ``BPMF/clib.py`` (clib.py:12-15).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "csrc", "_obj")
LIBNAME = "libbpmf_b200.so"
FEATURES_LIBNAME = "libbpmf_b200_features.so"  # detection traces (cuFFT + CUB); never loaded by the hot path
FEATURES_SOURCES = ["features.cu"]
SOURCES = ["abi.cu", "mf_prepass.cu", "mf_fp32.cu", "mf_tc.cu", "beam.cu", "libc_gpu.cu", "intertp.cu", "moveouts.cu", "peaks.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def features_lib_path():
    return os.path.join(LIBDIR, FEATURES_LIBNAME)


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "bpmf_b200.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "bpmf_b200_features.h"))
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES + FEATURES_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, log))
        if verbose:
            print(log)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    out = lib_path()
    if force or jobs or _stale(out, objs):
        # link to a temporary name and rename: a snapshot of the tree never sees a half-written library
        cmd = [nvcc, "-shared", "-o", out + ".tmp"] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        os.replace(out + ".tmp", out)
    fobjs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in FEATURES_SOURCES]
    fout = features_lib_path()
    if force or jobs or _stale(fout, fobjs):
        cmd = [nvcc, "-shared", "-o", fout + ".tmp"] + fobjs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lcufft"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        os.replace(fout + ".tmp", fout)
    link = os.path.join(LIBDIR, "libc.so")
    if not os.path.islink(link) and not os.path.exists(link):
        os.symlink(LIBNAME, link)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
