"""Build recipe for the oracle (test infrastructure).

* ``oracle/liboracle.so``   -- our C99+OpenMP restatement of the two un-vendored kernels
  (mf_oracle.c, beam_oracle.c).
* ``oracle/_ref/libc.so``   -- the reference's OWN ``BPMF/libc.c`` compiled from where it lies under
  /root/reference (never copied into this repo).  Flags follow the reference Makefile:20,24-25
  (``-O3 -fopenmp -fPIC -ftree-vectorize -std=c99 -shared``) except ``-march=native`` which is
  replaced by ``-march=x86-64-v3`` because the artefact is built in the CPU container and then
  travels to a GPU box whose host CPU may differ.

Both artefacts are git-ignored but travel with gpurun snapshots.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIBC = "/root/reference/BPMF/libc.c"
CFLAGS = ["-O3", "-fopenmp", "-fPIC", "-ftree-vectorize", "-march=x86-64-v3", "-std=c99", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False, verbose=False):
    srcs = [os.path.join(HERE, "mf_oracle.c"), os.path.join(HERE, "beam_oracle.c")]
    out = os.path.join(HERE, "liboracle.so")
    if not force and _newer(out, srcs):
        return out
    cmd = ["gcc"] + CFLAGS + srcs + ["-o", out, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_oracle_native(verbose=False):
    """The same two restatements compiled ON THIS HOST with the reference's own flags, ``-march=native`` included
    (Makefile:20,24-25) -- what bench.py times as the CPU baseline on the GPU box.  The artefact is tagged with
    the CPU model and never reused on another host.  Returns the path, or None when gcc is missing."""
    import hashlib

    srcs = [os.path.join(HERE, "mf_oracle.c"), os.path.join(HERE, "beam_oracle.c")]
    feats = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                feats = ln
                break
    except OSError:
        pass
    tag = hashlib.sha1((cpu_model() + feats + os.uname().nodename).encode()).hexdigest()[:10]
    out_dir = os.path.join(HERE, "_native")
    out = os.path.join(out_dir, "liboracle-%s.so" % tag)
    if _newer(out, srcs):
        return out
    os.makedirs(out_dir, exist_ok=True)
    flags = [f if f != "-march=x86-64-v3" else "-march=native" for f in CFLAGS]
    cmd = ["gcc"] + flags + srcs + ["-o", out, "-lm"]
    if verbose:
        print(" ".join(cmd))
    try:
        subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        return None
    return out


def build_ref(force=False, verbose=False):
    """Compile the reference's libc.c -> oracle/_ref/libc.so (only where /root/reference exists)."""
    out_dir = os.path.join(HERE, "_ref", "lib")
    out = os.path.join(out_dir, "libc.so")
    if not os.path.exists(REF_LIBC):
        return out if os.path.exists(out) else None
    if not force and _newer(out, [REF_LIBC]):
        return out
    os.makedirs(out_dir, exist_ok=True)
    cmd = ["gcc"] + CFLAGS + [REF_LIBC, "-o", out, "-lm"]
    if verbose:
        print(" ".join(cmd))
    # the reference prints %d for size_t: warnings are expected, errors are not
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    return out


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv, verbose=True))
    print(build_ref(force="--force" in sys.argv, verbose=True))
