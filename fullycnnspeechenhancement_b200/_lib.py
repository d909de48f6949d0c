"""ctypes binding of librced_b200.so (C ABI declared in include/rced.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/build.sh``.  There is
no fallback: if the shared object is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librced_b200.so")
ABI_VERSION = 3
VARIANT_FFMA = 0   # FP32 FFMA network kernel
VARIANT_TC = 1     # tcgen05 tensor-core kernel (FP16 x3 split)

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/rced.h declaration by declaration
SIGNATURES = {
    "rced_abi_version": (ctypes.c_int, []),
    "rced_last_error": (ctypes.c_char_p, []),
    "rced_num_frames": (c_i64, [c_i64]),
    "rced_folded_weight_count": (c_i64, [ctypes.c_int]),
    "rced_num_layers": (ctypes.c_int, [ctypes.c_int]),
    "rced_layer_shape": (ctypes.c_int, [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 4),
    "rced_packed_weight_count": (c_i64, [ctypes.c_int]),
    "rced_pack_weights": (ctypes.c_int, [ctypes.c_int, c_p, ctypes.c_size_t, c_p, ctypes.c_size_t]),
    "rced_debug_layout": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_i64), ctypes.c_int]),
    "rced_mac_per_frame": (c_i64, [ctypes.c_int, ctypes.c_int]),
    "rced_create": (ctypes.c_int, [ctypes.c_int, c_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_p)]),
    "rced_destroy": (None, [c_p]),
    "rced_arch": (ctypes.c_int, [c_p]),
    "rced_device": (ctypes.c_int, [c_p]),
    "rced_set_skip_in_tmem": (ctypes.c_int, [c_p, ctypes.c_int]),
    "rced_set_variant": (ctypes.c_int, [c_p, ctypes.c_int]),
    "rced_variant": (ctypes.c_int, [c_p]),
    "rced_tc_status": (ctypes.c_int, [c_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint)]),
    "rced_tc_image_bytes": (c_i64, [ctypes.c_int]),
    "rced_tc_bias_count": (c_i64, [ctypes.c_int]),
    "rced_tc_pack_weights": (ctypes.c_int, [ctypes.c_int, c_p, ctypes.c_size_t, c_p, ctypes.c_size_t, c_p, ctypes.c_size_t]),
    "rced_tc_layout": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_i64), ctypes.c_int]),
    "rced_stft": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, ctypes.c_int, c_i64, c_p, c_p, c_p]),
    "rced_forward": (ctypes.c_int, [c_p, c_p, c_p, ctypes.c_int, c_i64, c_p, c_p]),
    "rced_istft": (ctypes.c_int, [c_p, c_p, c_p, c_p, ctypes.c_int, c_i64, ctypes.c_int, c_p, c_p, c_p, c_p]),
    "rced_enhance": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, ctypes.c_int, c_i64, c_i64, ctypes.c_int,
                                    c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "rced_enhance_host": (ctypes.c_int, [c_p, c_p, c_p, c_p, ctypes.c_int, ctypes.c_int, c_p, c_p, c_p]),
    "rced_enhance_host_async": (ctypes.c_int, [c_p, c_p, c_p, c_p, ctypes.c_int, ctypes.c_int, c_p, c_p, c_p]),
    "rced_host_sync": (ctypes.c_int, [c_p]),
    "rced_host_config": (ctypes.c_int, [c_p, c_i64, c_i64]),
    "rced_host_set_relay": (ctypes.c_int, [c_p, ctypes.c_int]),
    "rced_host_link_probe": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_double)]),
    "rced_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_p)]),
    "rced_host_free": (ctypes.c_int, [c_p]),
    "rced_mag_phase": (ctypes.c_int, [ctypes.c_int, c_p, c_i64, c_p, c_p, c_p]),
    "rced_sdr_sums": (ctypes.c_int, [ctypes.c_int, c_p, c_p, c_p, c_p, c_p, ctypes.c_int, c_i64, c_p, c_p]),
    "rced_ffma_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "rced_selftest_tmem": (ctypes.c_int, [ctypes.c_int]),
    "rced_launch_count": (c_i64, []),
}


ERR_ARG, ERR_CUDA, ERR_STATE = -1, -2, -3


class RcedError(RuntimeError):
    """`code` is the library's return value (ERR_ARG / ERR_CUDA / ERR_STATE), None for loader errors."""

    def __init__(self, message, code=None):
        RuntimeError.__init__(self, message)
        self.code = code


_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RcedError(
                "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or fullycnnspeechenhancement_b200/csrc/build.sh). There is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        if handle.rced_abi_version() != ABI_VERSION:
            raise RcedError("librced_b200.so ABI version %d, python expects %d" % (handle.rced_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(code):
    if code != 0:
        raise RcedError("librced_b200 error %d: %s" % (code, lib().rced_last_error().decode("utf-8", "replace")), code)


def num_frames(n_samples):
    """T(L) of data_utils/audio_feature.py:67-70, from the library."""
    return int(lib().rced_num_frames(int(n_samples)))
