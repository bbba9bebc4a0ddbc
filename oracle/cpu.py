"""ctypes loader for the C++ CPU restatement (oracle/cpu_prover.cpp).

TEST INFRASTRUCTURE ONLY (see oracle/bn254.py header for the import rule).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "liboracle_cpu.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_hw_threads.restype = C.c_int
        _lib.oracle_groth16_prove.restype = C.c_int
        _lib.oracle_msm_g1.restype = C.c_double
    return _lib


def hw_threads() -> int:
    return lib().oracle_hw_threads()


def prove(params: bytes, n_gates: int, n_in: int, n_aux: int, rowptr, col, coef, inputs, aux, r, s,
          nthreads: int, want_h: bool = False):
    """rowptr/col: 3 uint32 arrays; coef: 3 uint64[nnz,4] Montgomery; inputs/aux/r/s uint64 Montgomery.
    Returns (proof_raw bytes, h or None, stage seconds [eval, fft, multiexp, total])."""
    L = lib()
    rp = [np.ascontiguousarray(x, dtype=np.uint32) for x in rowptr]
    cl = [np.ascontiguousarray(x, dtype=np.uint32) for x in col]
    cf = [np.ascontiguousarray(x, dtype=np.uint64) for x in coef]
    arr = lambda xs: (C.c_void_p * 3)(*[x.ctypes.data for x in xs])
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    aux = np.ascontiguousarray(aux, dtype=np.uint64)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    proof = np.zeros(256, dtype=np.uint8)
    m = 1
    while m < n_gates + n_in:
        m *= 2
    m = max(m, 2)
    h = np.zeros((m - 1, 4), dtype=np.uint64) if want_h else None
    st = (C.c_double * 4)()
    pbuf = np.frombuffer(params, dtype=np.uint8) if isinstance(params, (bytes, bytearray)) else np.ascontiguousarray(params, dtype=np.uint8)
    rc = L.oracle_groth16_prove(C.c_void_p(pbuf.ctypes.data), C.c_size_t(pbuf.size), C.c_uint32(n_gates), C.c_uint32(n_in),
                                C.c_uint32(n_aux), arr(rp), arr(cl), arr(cf), C.c_void_p(inputs.ctypes.data),
                                C.c_void_p(aux.ctypes.data), C.c_void_p(r.ctypes.data), C.c_void_p(s.ctypes.data),
                                C.c_int(nthreads), C.c_void_p(proof.ctypes.data),
                                C.c_void_p(h.ctypes.data) if h is not None else None, st)
    if rc != 0:
        raise RuntimeError(f"oracle_groth16_prove failed: {rc}")
    return proof.tobytes(), h, list(st)


def msm_g1(bases_raw: np.ndarray, scalars: np.ndarray, nthreads: int):
    L = lib()
    n = scalars.shape[0]
    out = np.zeros(64, dtype=np.uint8)
    sec = L.oracle_msm_g1(C.c_void_p(bases_raw.ctypes.data), C.c_void_p(scalars.ctypes.data), C.c_size_t(n),
                          C.c_int(nthreads), C.c_void_p(out.ctypes.data))
    return out.tobytes(), sec


def csr_from_gates(gates, n_in: int):
    """Python gate list (oracle.groth16 format) -> CSR arrays for prove()."""
    from . import codec
    rowptr = [[0], [0], [0]]
    col = [[], [], []]
    coef = [[], [], []]
    for g in gates:
        for m in range(3):
            for c, (tag, idx) in g[m]:
                col[m].append(idx if tag == 0 else n_in + idx)
                coef[m].append(codec.fr_raw(c))
            rowptr[m].append(len(col[m]))
    return ([np.array(x, dtype=np.uint32) for x in rowptr], [np.array(x, dtype=np.uint32) for x in col],
            [np.frombuffer(b"".join(x), dtype=np.uint64).reshape(-1, 4).copy() if x else np.zeros((0, 4), np.uint64)
             for x in coef])
