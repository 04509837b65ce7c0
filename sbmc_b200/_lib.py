"""ctypes binding of libsbmc_b200.so (C ABI declared in include/sbmc_b200.h).

PyTorch is only used by the callers for device memory and streams; the library
itself takes raw pointers.  There is no CPU or PyTorch fallback: if the library
is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsbmc_b200.so")

_i64, _int, _ptr, _f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_float

# name -> (restype, argtypes); mirrors include/sbmc_b200.h one to one.
SIGNATURES = {
    "sbmc_b200_version": (_int, []),
    "sbmc_b200_last_error": (ctypes.c_char_p, []),
    "sbmc_b200_force_generic": (_int, [_int]),
    "sbmc_b200_last_path": (_int, []),
    "sbmc_b200_launch_count": (_i64, []),
    "sbmc_b200_timing_enable": (_int, [_int]),
    "sbmc_b200_timing_collect": (_int, [_ptr, _ptr]),
    "sbmc_scatter2gather_f32":
        (_int, [_ptr, _ptr, _i64, _int, _int, _i64, _i64, _ptr]),
    "sbmc_kernel_weighting_fwd_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64, _int, _int, _ptr]),
    "sbmc_kernel_weighting_bwd_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64,
                _int, _int, _ptr]),
    "sbmc_progressive_splat_fwd_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64, _int, _int, _int,
                _int, _ptr]),
    "sbmc_progressive_splat_bwd_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64, _int, _int, _ptr]),
    "sbmc_conv1x1_chain_f32":
        (_int, [_ptr, _int, _i64, _ptr, _int, _i64, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                _int, _int, _int, _int, _ptr, _i64, _i64, _i64, _ptr]),
    "sbmc_conv1x1_chain_nhwc_bf16":
        (_int, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _int, _int,
                _int, _ptr, _i64, _int, _i64, _i64, _ptr]),
    "sbmc_chain_samples_nhwc_bf16":
        (_int, [_ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr,
                _int, _int, _int, _int, _ptr, _i64, _i64, _ptr, _i64, _int, _i64, _i64, _i64,
                _i64, _ptr]),
    "sbmc_b200_conv3x3_pair": (_int, [_int]),
    "sbmc_b200_conv3x3_linear": (_int, [_int]),
    "sbmc_conv3x3_nhwc_bf16":
        (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _int, _int, _ptr]),
    "sbmc_upsample_concat_nhwc_bf16":
        (_int, [_ptr, _ptr, _ptr, _i64, _int, _int, _int, _int, _int, _int, _ptr]),
    "sbmc_linear_nhwc_bf16":
        (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _int, _ptr]),
    "sbmc_linear2_nhwc_bf16":
        (_int, [_ptr, _int, _ptr, _int, _i64, _i64, _ptr, _ptr, _ptr, _int, _ptr, _int, _i64,
                _i64, _int, _i64, _int, _int, _ptr]),
    "sbmc_wgrad_nhwc_bf16":
        (_int, [_ptr, _ptr, _i64, _i64, _int, _int, _int, _ptr, _ptr, _i64, _int, _int, _ptr,
                _ptr]),
    "sbmc_wgrad3x3_nhwc_bf16":
        (_int, [_ptr, _ptr, _i64, _int, _int, _int, _int, _int, _ptr, _ptr, _ptr, _ptr]),
    "sbmc_weight_bank_run": (_int, [_ptr, _ptr, _i64, _int, _ptr]),
    "sbmc_conv3x3_masked_nhwc_bf16":
        (_int, [_ptr, _ptr, _ptr, _ptr, _int, _ptr, _i64, _int, _int, _int, _int, _int, _ptr]),
    "sbmc_spp_reduce_nhwc_bf16": (_int, [_ptr, _ptr, _int, _i64, _int, _i64, _int, _f32, _ptr]),
    "sbmc_bcast_add_nhwc_bf16": (_int, [_ptr, _ptr, _ptr, _i64, _int, _i64, _int, _f32, _ptr]),
    "sbmc_maxpool2x2_bwd_nhwc_bf16":
        (_int, [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _int, _int, _int, _int, _ptr]),
    "sbmc_upsample_bwd_nhwc_bf16":
        (_int, [_ptr, _i64, _ptr, _ptr, _i64, _int, _int, _int, _int, _int, _int, _ptr]),
    "sbmc_dact_bf16": (_int, [_ptr, _ptr, _ptr, _i64, _int, _ptr]),
    "sbmc_colsum_bf16": (_int, [_ptr, _i64, _i64, _int, _ptr, _int, _ptr, _ptr]),
    "sbmc_maxpool2x2_nhwc_bf16": (_int, [_ptr, _ptr, _i64, _int, _int, _int, _ptr]),
    "sbmc_bias_act_nhwc_bf16": (_int, [_ptr, _ptr, _i64, _int, _int, _ptr]),
    "sbmc_nchw_to_nhwc_bf16": (_int, [_ptr, _i64, _ptr, _i64, _i64, _int, _i64, _int, _ptr]),
    "sbmc_lz4_frames_inflate": (_int, [_ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "sbmc_tile_assemble_f32":
        (_int, [_ptr, _ptr, _i64, _i64, _int, _int, _int, _int, _int, _int, _ptr, _ptr, _ptr,
                _ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr]),
    "sbmc_multi_tensor_grad_norm_f32":
        (_int, [_ptr, _ptr, _i64, _ptr, ctypes.c_float, _ptr, _ptr]),
    "sbmc_multi_tensor_adam_f32":
        (_int, [_ptr, _ptr, _i64, _ptr] + [ctypes.c_double] * 6 + [_ptr]),
    "sbmc_multi_tensor_adam_devstep_f32":
        (_int, [_ptr, _ptr, _i64, _ptr] + [ctypes.c_double] * 4 + [_ptr, _ptr]),
    "sbmc_kernel_weighting_fwd_band_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64, _int, _int, _int,
                _int, _ptr]),
    "sbmc_kernel_weighting_bwd_band_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64, _int,
                _int, _int, _int, _ptr]),
    "sbmc_scatter2gather_host_f32":
        (_int, [_ptr, _ptr, _i64, _int, _int, _i64, _i64, _int]),
    "sbmc_kernel_weighting_fwd_host_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64, _int, _int, _int]),
    "sbmc_kernel_weighting_bwd_host_f32":
        (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _i64, _i64,
                _int, _int, _int]),
    "sbmc_b200_host_release": (_int, []),
}

_lib = None


class SbmcB200Error(RuntimeError):
    """A libsbmc_b200 call returned a non-zero status."""


def load():
    """Load the shared library (building it with nvcc first if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    # (re)build when the library is absent OR older than csrc/ / include/ (a stale .so
    # loaded against the current ctypes SIGNATURES could corrupt memory silently);
    # build() is a no-op when nothing changed.  Without nvcc an existing library is
    # used as it is -- the GPU box runs the prebuilt file.
    from . import build as _build
    if not os.path.exists(LIB_PATH):
        _build.build()
    elif _build.nvcc_available():
        try:
            _build.build()
        except RuntimeError as e:
            raise SbmcB200Error("libsbmc_b200.so is older than its sources and the rebuild "
                                "failed: %s" % e)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise SbmcB200Error(
            "libsbmc_b200.so could not be loaded (%s); there is no fallback "
            "path -- build it with `python -m sbmc_b200.build`" % e)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EUNSUPPORTED = -5


def check(rc, what):
    if rc != 0:
        msg = load().sbmc_b200_last_error()
        raise SbmcB200Error("%s failed (code %d): %s" % (
            what, rc, msg.decode("utf-8", "replace") if msg else ""))


def launch_count():
    return int(load().sbmc_b200_launch_count())


def force_generic(flag):
    return int(load().sbmc_b200_force_generic(1 if flag else 0))


def last_path():
    return int(load().sbmc_b200_last_path())


KERNEL_KINDS = {0: "kw_fwd", 1: "kw_bwd_dweights", 2: "kw_bwd_ddata", 3: "s2g",
                4: "other", 5: "splat_fwd", 6: "splat_bwd", 7: "conv1x1_chain", 8: "tiles", 9: "optim", 10: "conv3x3"}
NUM_KERNEL_KINDS = 11


def timing_enable(flag):
    return int(load().sbmc_b200_timing_enable(1 if flag else 0))


def timing_collect():
    """{kernel kind: (device ms, launches)} since the previous collect."""
    ms = (ctypes.c_double * NUM_KERNEL_KINDS)()
    cnt = (ctypes.c_int64 * NUM_KERNEL_KINDS)()
    check(load().sbmc_b200_timing_collect(ctypes.cast(ms, _ptr), ctypes.cast(cnt, _ptr)),
          "timing_collect")
    return {KERNEL_KINDS.get(k, "kind%d" % k): (ms[k], int(cnt[k]))
            for k in range(NUM_KERNEL_KINDS) if cnt[k]}
