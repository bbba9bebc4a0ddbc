"""ctypes binding of libfawkes_b200.so (the C ABI in include/fawkes_b200.h).

This is the only route to the kernels: there is no Python or CPU implementation of the
proving path in this package.  If the shared library is missing, importing this module
raises; if no CUDA device is present, fb_init fails with FB_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfawkes_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C fawkes-crypto_b200/csrc).  There is no fallback implementation.")

lib = C.CDLL(LIB_PATH)

u8p, u64p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
vp = C.c_void_p


class PkInfo(C.Structure):
    _fields_ = [("n_in", C.c_uint32), ("n_aux", C.c_uint32), ("n_gates", C.c_uint32),
                ("log_m", C.c_uint32), ("len_h", C.c_uint32), ("len_l", C.c_uint32),
                ("len_a", C.c_uint32), ("len_b", C.c_uint32), ("nnz", C.c_uint64),
                ("hbm_bytes", C.c_uint64), ("g1_digit_slots", C.c_uint64), ("g2_digit_slots", C.c_uint64),
                ("msm_window_bits", C.c_uint32), ("msm_windows", C.c_uint32), ("msm_tables", C.c_uint32),
                ("reserved0", C.c_uint32), ("table_bytes", C.c_uint64)]


# every symbol declared in include/fawkes_b200.h: (restype, argtypes)
SIGNATURES = {
    "fb_init": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
    "fb_shutdown": (None, [vp]),
    "fb_last_error": (C.c_char_p, []),
    "fb_device_count": (C.c_int, []),
    "fb_free": (None, [vp]),
    "fb_circuit_from_gates": (C.c_int, [vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "fb_circuit_from_raw_gates": (C.c_int, [vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "fb_circuit_from_gates_gpu": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), f32p]),
    "fb_circuit_from_raw_gates_gpu": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), f32p]),
    "fb_circuit_free": (None, [vp]),
    "fb_circuit_shape": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "fb_pk_load": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, C.c_int, C.POINTER(vp)]),
    "fb_pk_load_circuit": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_int, C.POINTER(vp)]),
    "fb_pk_load_shard": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "fb_pk_free": (None, [vp]),
    "fb_pk_get_info": (C.c_int, [vp, C.POINTER(PkInfo)]),
    "fb_prove": (C.c_int, [vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp]),
    "fb_prove_batch": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp]),
    "fb_stream_open": (C.c_int, [vp, vp, C.c_int, C.POINTER(vp)]),
    "fb_stream_submit": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint64)]),
    "fb_stream_wait": (C.c_int, [vp, C.c_uint64, vp]),
    "fb_stream_close": (None, [vp]),
    "fb_prove_device": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "fb_prove_partial": (C.c_int, [vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp]),
    "fb_prove_finish": (C.c_int, [vp, C.c_size_t, vp, C.c_int, vp, vp, vp]),
    "fb_prove_timings": (C.c_int, [vp, f32p]),
    "fb_setup": (C.c_int, [vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "fb_setup_shard": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "fb_verify": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint32, C.POINTER(C.c_int)]),
    "fb_circuit_synth": (C.c_int, [C.c_uint64, C.c_uint64, C.POINTER(vp)]),
    "fb_circuit_witness": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "fb_synth_trapdoor": (C.c_int, [C.c_uint64, vp]),
    "fb_circuit_csr": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "fb_launch_count": (C.c_uint64, []),
    "fb_set_serial": (None, [C.c_int]),
    "fb_set_msm_tables": (None, [C.c_int]),
    "fb_set_prove_graph": (None, [C.c_int]),
    "fb_kernel_stats_enable": (None, [C.c_int]),
    "fb_kernel_stats_reset": (None, []),
    "fb_kernel_stats": (C.c_int, [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "fb_test_field": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_uint64]),
    "fb_test_ntt": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "fb_test_h": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, f32p]),
    "fb_test_dist_h": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "fb_dist_unique_id": (C.c_int, [vp]),
    "fb_dist_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "fb_test_msm": (C.c_int, [vp, C.c_int, vp, vp, C.c_uint64, vp, C.c_int, f32p]),
    "fb_test_msm_plan": (C.c_int, [C.c_uint32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fb_test_pairing": (C.c_int, [C.c_uint64, C.c_int, C.POINTER(C.c_int)]),
    "fb_test_fixed_base": (C.c_int, [vp, C.c_int, vp, C.c_uint64, vp]),
    "fb_probe_imad": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "fb_probe_rate": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "fb_probe_fr_mul": (C.c_int, [vp, C.POINTER(C.c_double)]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

FB_LOAD_CHECKED, FB_LOAD_NO_INFINITY = 1, 2

ERR_NAMES = {0: "FB_OK", -1: "FB_ERR_ARG", -2: "FB_ERR_CUDA", -3: "FB_ERR_FORMAT", -4: "FB_ERR_DOMAIN",
             -5: "FB_ERR_IDENTITY", -6: "FB_ERR_DENSITY", -7: "FB_ERR_VK", -8: "FB_ERR_HOST"}


class FbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def last_error() -> str:
    return (lib.fb_last_error() or b"").decode(errors="replace")


def check(rc: int):
    if rc != 0:
        raise FbError(rc, last_error())


def ptr(a):
    """Address of a bytes / numpy buffer (the caller keeps the object alive)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if isinstance(a, bytes):
        return C.cast(C.c_char_p(a), vp).value
    if isinstance(a, int):
        return a
    raise TypeError(type(a))
