"""CPU oracle for the SBMC kernel-splatting hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``sbmc_b200``) never does; it fails loudly when its CUDA library is missing.

Python face of ``sbmc_oracle.c`` (see that file's header for the reference
file:line each function restates and for the parity-pin statement).  Works on
torch CPU tensors (float32, contiguous), mirroring the six entry points the
reference binds in ``sbmc.halide_ops`` (setup.py:65-84 of the reference): same
argument order, caller-allocated outputs.
"""
import ctypes
import hashlib
import os
import subprocess

import torch as th

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsbmc_oracle.so")
_SIG_PATH = os.path.join(_HERE, ".build_sig")
_lib = None


def _cpu_signature():
    """-march=native code must not travel between hosts: key the build on the
    CPU flag set and rebuild when it changes (gpurun boxes differ from the
    build container)."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as fid:
            for line in fid:
                if line.startswith("flags"):
                    flags = " ".join(sorted(line.split(":", 1)[1].split()))
                    break
    except OSError:
        pass
    src = b"".join(open(os.path.join(_HERE, f), "rb").read()
                   for f in ("sbmc_oracle.c", "lz4_oracle.c", "Makefile"))
    return hashlib.sha1(flags.encode() + src).hexdigest()


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    sig = _cpu_signature()
    have = None
    if os.path.exists(_SIG_PATH):
        have = open(_SIG_PATH).read().strip()
    if force or have != sig or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE, "clean", "all"])
        with open(_SIG_PATH, "w") as fid:
            fid.write(sig)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, fp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        L.sbmc_oracle_kernel_weighting_f32.argtypes = [
            fp, fp, fp, fp, i64, i32, i64, i64, i32, i32]
        L.sbmc_oracle_kernel_weighting_grad_f32.argtypes = [
            fp, fp, fp, fp, fp, fp, fp, i64, i32, i64, i64, i32, i32]
        L.sbmc_oracle_scatter2gather_f32.argtypes = [
            fp, fp, i64, i32, i32, i64, i64]
        L.sbmc_oracle_set_num_threads.argtypes = [i32]
        for f in (L.sbmc_oracle_kernel_weighting_f32,
                  L.sbmc_oracle_kernel_weighting_grad_f32,
                  L.sbmc_oracle_scatter2gather_f32,
                  L.sbmc_oracle_num_threads):
            f.restype = i32
        L.sbmc_oracle_lz4_frame_decompress.argtypes = [fp, i64, fp, i64, fp]
        L.sbmc_oracle_lz4_frame_decompress.restype = i32
        L.sbmc_oracle_xxh32.argtypes = [fp, i64, ctypes.c_uint32]
        L.sbmc_oracle_xxh32.restype = ctypes.c_uint32
        _lib = L
    return _lib


def num_threads():
    return lib().sbmc_oracle_num_threads()


def set_num_threads(n):
    lib().sbmc_oracle_set_num_threads(int(n))


def _chk(*tensors):
    for t in tensors:
        if t.device.type != "cpu" or t.dtype != th.float32 or not t.is_contiguous():
            raise ValueError("oracle wants contiguous float32 CPU tensors")


def _rc(rc, name):
    if rc != 0:
        raise RuntimeError("%s failed with code %d" % (name, rc))


# -- the six reference entry points (CPU flavour) ------------------------------
def kernel_weighting_cpu_float32(data, weights, output, sum_w):
    """reference: ops.kernel_weighting_cpu_float32 (sbmc/functions.py:97-98)."""
    _chk(data, weights, output, sum_w)
    n, c, h, w = data.shape
    _, kh, kw, _, _ = weights.shape
    _rc(lib().sbmc_oracle_kernel_weighting_f32(
        data.data_ptr(), weights.data_ptr(), output.data_ptr(),
        sum_w.data_ptr(), n, c, h, w, kh, kw), "kernel_weighting")
    return 0


def kernel_weighting_grad_cpu_float32(data, weights, sum_w, d_output, d_sum_w,
                                      d_data, d_weights):
    """reference: ops.kernel_weighting_grad_cpu_float32 (sbmc/functions.py:113-114)."""
    _chk(data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
    n, c, h, w = data.shape
    _, kh, kw, _, _ = weights.shape
    _rc(lib().sbmc_oracle_kernel_weighting_grad_f32(
        data.data_ptr(), weights.data_ptr(), sum_w.data_ptr(),
        d_output.data_ptr(), d_sum_w.data_ptr(), d_data.data_ptr(),
        d_weights.data_ptr(), n, c, h, w, kh, kw), "kernel_weighting_grad")
    return 0


def scatter2gather_cpu_float32(weights, output):
    """reference: ops.scatter2gather_cpu_float32 (sbmc/functions.py:58-59)."""
    _chk(weights, output)
    n, kh, kw, h, w = weights.shape
    _rc(lib().sbmc_oracle_scatter2gather_f32(
        weights.data_ptr(), output.data_ptr(), n, kh, kw, h, w),
        "scatter2gather")
    return 0


# -- functional conveniences ---------------------------------------------------
def kernel_weighting(data, weights):
    data = data.detach().cpu().float().contiguous()
    weights = weights.detach().cpu().float().contiguous()
    n, c, h, w = data.shape
    out = th.empty_like(data)
    sum_w = th.empty(n, h, w, dtype=th.float32)
    kernel_weighting_cpu_float32(data, weights, out, sum_w)
    return out, sum_w


def kernel_weighting_grad(data, weights, d_output, d_sum_w):
    data = data.detach().cpu().float().contiguous()
    weights = weights.detach().cpu().float().contiguous()
    d_output = d_output.detach().cpu().float().contiguous()
    d_sum_w = d_sum_w.detach().cpu().float().contiguous()
    n, c, h, w = data.shape
    d_data = th.empty_like(data)
    d_weights = th.empty_like(weights)
    sum_w = th.zeros(n, h, w, dtype=th.float32)  # unread by the pipeline
    kernel_weighting_grad_cpu_float32(data, weights, sum_w, d_output, d_sum_w,
                                      d_data, d_weights)
    return d_data, d_weights


def scatter2gather(weights):
    weights = weights.detach().cpu().float().contiguous()
    out = th.empty_like(weights)
    scatter2gather_cpu_float32(weights, out)
    return out


# -- the tile reader's decompressor (reference: lz4.frame.decompress) ---------------
class Lz4Error(RuntimeError):
    def __init__(self, code):
        RuntimeError.__init__(self, "lz4 frame decode failed with code %d" % code)
        self.code = code


def lz4_frame_decompress(buf, max_size=None):
    """bytes -> bytes, like `lz4.frame.decompress` (sbmc/datasets.py:578).  The
    output buffer grows until the frame fits (the format does not have to carry
    its content size)."""
    buf = bytes(buf)
    cap = max(4 * len(buf), 1 << 16) if max_size is None else int(max_size)
    src = ctypes.create_string_buffer(buf, len(buf))
    while True:
        dst = ctypes.create_string_buffer(cap)
        out_len = ctypes.c_int64(0)
        rc = lib().sbmc_oracle_lz4_frame_decompress(
            ctypes.cast(src, ctypes.c_void_p), len(buf), ctypes.cast(dst, ctypes.c_void_p),
            cap, ctypes.cast(ctypes.pointer(out_len), ctypes.c_void_p))
        if rc == 4 and max_size is None and cap < (1 << 34):   # overflow: retry larger
            cap *= 4
            continue
        if rc != 0:
            raise Lz4Error(rc)
        return dst.raw[:out_len.value]


def xxh32(buf, seed=0):
    buf = bytes(buf)
    src = ctypes.create_string_buffer(buf, len(buf))
    return int(lib().sbmc_oracle_xxh32(ctypes.cast(src, ctypes.c_void_p), len(buf), seed))
