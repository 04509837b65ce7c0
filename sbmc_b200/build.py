"""Builds libsbmc_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch involvement: the library is plain CUDA behind ``extern "C"`` entry
points (include/sbmc_b200.h).  nvcc cross-compiles without a GPU.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libsbmc_b200.so")
SOURCES = ["runtime.cu", "generic.cu", "kw_launch.cu", "s2g.cu", "capi.cu",
           "host_stream.cu", "splat.cu", "splat_bwd.cu", "conv1x1.cu", "chain_v3.cu", "conv3x3.cu", "linear.cu", "wgrad.cu", "train_ops.cu", "weight_bank.cu", "unet_ops.cu", "tiles.cu", "optim.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
    "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas=-v",
]


def _extra_flags():
    """SBMC_B200_NVCC_FLAGS="-DSBMC_LZ4_WIDE_COPY ..." adds defines for A/B runs of
    experimental paths (none is part of the default build)."""
    return os.environ.get("SBMC_B200_NVCC_FLAGS", "").split()


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libsbmc_b200.so")
    return cand


def nvcc_available():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return os.path.exists(cand)


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES
            if os.path.exists(os.path.join(CSRC, s))]


HASH_PATH = os.path.join(HERE, "build", "source.sha1")


def _source_hash():
    """Content hash of everything the library is built from (sources, headers, flags):
    unlike mtimes it survives the copy onto the GPU box."""
    h = hashlib.sha1()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    paths += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for path in paths:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fid:
            h.update(fid.read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags() + SOURCES).encode())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fid:
        return fid.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library; returns its path.
    Processes that race here (one rank per GPU) serialise on a file lock and all but
    the first find the library up to date."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in _sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + _extra_flags() + ["-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(
            cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    logs = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        logs.append("== %s\n%s" % (os.path.basename(src), out))
        failed |= p.returncode != 0
    log = "\n".join(logs)
    with open(os.path.join(HERE, "build", "nvcc.log"), "w") as fid:
        fid.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libsbmc_b200.so")
    if verbose:
        print(log)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + [
        "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
        "-Xlinker", "--exclude-libs,ALL"]
    subprocess.check_call(cmd)
    with open(HASH_PATH, "w") as fid:
        fid.write(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
