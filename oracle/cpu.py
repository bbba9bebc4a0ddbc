"""ctypes loader for the C++ CPU restatement (oracle/cpu_prover.cpp, oracle/cpu_setup.cpp).

TEST INFRASTRUCTURE ONLY (see oracle/bn254.py header for the import rule).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "liboracle_cpu.so")
_lib = None
_lib_kind = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _native_lib():
    """-march=native build for THIS host (the CPU baseline is compiled for the cores it
    is timed on).  Built once per CPU model into oracle/_ref/native-<hash>/; None when that is not possible."""
    if os.environ.get("FB_ORACLE_NATIVE", "1") == "0":
        return None
    try:
        flags = ""
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags") or line.startswith("model name"):
                    flags += line
                    if line.startswith("flags"):
                        break
        srcs = b"".join(open(os.path.join(_HERE, n), "rb").read() for n in ("cpu_common.h", "cpu_prover.cpp", "cpu_setup.cpp"))
        tag = hashlib.sha256(flags.encode() + srcs).hexdigest()[:12]
        out = os.path.join(_HERE, "_ref", f"native-{tag}")
        so = os.path.join(out, "liboracle_cpu.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", _HERE, "native", f"OUT={out}"], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL, timeout=300)
        return so if os.path.exists(so) else None
    except Exception:
        return None


def lib():
    global _lib, _lib_kind
    if _lib is None:
        path = _native_lib()
        _lib_kind = "g++ -O3 -march=native"
        if path is None:
            if not os.path.exists(LIB):
                build()
            path, _lib_kind = LIB, "g++ -O3 -march=x86-64-v3"
        _lib = C.CDLL(path)
        _lib.oracle_hw_threads.restype = C.c_int
        _lib.oracle_groth16_prove.restype = C.c_int
        _lib.oracle_msm_g1.restype = C.c_double
        _lib.oracle_circuit_synth.restype = C.c_void_p
        _lib.oracle_circuit_from_csr.restype = C.c_void_p
        _lib.oracle_setup.restype = C.c_int
    return _lib


def build_kind() -> str:
    lib()
    return _lib_kind


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


class Circuit:
    """R1CS (+ witness for synthetic circuits) held by the C++ oracle; numpy views, no copies."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle circuit construction failed")
        self.handle = C.c_void_p(handle)
        L = lib()
        shape = (C.c_uint64 * 4)()
        rp, cl, cf = (C.c_void_p * 3)(), (C.c_void_p * 3)(), (C.c_void_p * 3)()
        nnz = (C.c_uint64 * 3)()
        pin, paux = C.c_void_p(), C.c_void_p()
        L.oracle_circuit_view(self.handle, shape, rp, cl, cf, nnz, C.byref(pin), C.byref(paux))
        self.n_in, self.n_aux, self.n_gates, self.nnz = (int(x) for x in shape)

        def view(ptr, n, ctype, cols=None):
            if not n:
                return np.zeros((0, cols) if cols else 0, dtype=np.dtype(ctype))
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n * (cols or 1),))
            return a.reshape(n, cols) if cols else a

        self.rowptr = [view(rp[m], self.n_gates + 1, C.c_uint32) for m in range(3)]
        self.col = [view(cl[m], int(nnz[m]), C.c_uint32) for m in range(3)]
        self.coef = [view(cf[m], int(nnz[m]), C.c_uint64, 4) for m in range(3)]
        self.inputs = view(pin, self.n_in, C.c_uint64, 4) if pin.value else None
        self.aux = view(paux, self.n_aux, C.c_uint64, 4) if paux.value else None

    @classmethod
    def synthetic(cls, n_rows: int, seed: int) -> "Circuit":
        return cls(lib().oracle_circuit_synth(C.c_uint64(n_rows), C.c_uint64(seed)))

    @classmethod
    def from_csr(cls, n_gates, n_in, n_aux, rowptr, col, coef) -> "Circuit":
        rp = [np.ascontiguousarray(x, dtype=np.uint32) for x in rowptr]
        cl = [np.ascontiguousarray(x, dtype=np.uint32) for x in col]
        cf = [np.ascontiguousarray(x, dtype=np.uint64) for x in coef]
        arr = lambda xs: (C.c_void_p * 3)(*[x.ctypes.data for x in xs])
        return cls(lib().oracle_circuit_from_csr(C.c_uint32(n_gates), C.c_uint32(n_in), C.c_uint32(n_aux), arr(rp), arr(cl),
                                                 arr(cf)))

    def close(self):
        if self.handle:
            self.rowptr = self.col = self.coef = self.inputs = self.aux = None
            lib().oracle_circuit_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_trapdoor(seed: int) -> np.ndarray:
    """alpha beta gamma delta tau r s as uint64[7,4] Montgomery limbs (stream seed ^ 0xB11D)."""
    out = np.zeros((7, 4), dtype=np.uint64)
    lib().oracle_synth_trapdoor(C.c_uint64(seed), C.c_void_p(out.ctypes.data))
    return out


class ParamsBuf:
    """bellman Parameters bytes produced by oracle_setup (malloc'ed by the library, freed with the object)."""

    def __init__(self, ptr, n):
        self._ptr, self.size = ptr, n
        self.array = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,))

    def tobytes(self) -> bytes:
        return self.array.tobytes()

    def __del__(self):
        try:
            if self._ptr:
                self.array = None
                lib().oracle_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def setup(circ: Circuit, trapdoor_mont: np.ndarray, nthreads: int):
    """Groth16 CRS for `circ` (bellman Parameters bytes).  Returns (ParamsBuf, [scalar_s, point_s, total_s])."""
    td = np.ascontiguousarray(trapdoor_mont[:5], dtype=np.uint64)
    out, n = C.c_void_p(), C.c_size_t()
    st = (C.c_double * 3)()
    rc = lib().oracle_setup(circ.handle, C.c_void_p(td.ctypes.data), C.c_int(nthreads), C.byref(out), C.byref(n), st)
    if rc != 0:
        raise RuntimeError(f"oracle_setup failed: {rc}")
    return ParamsBuf(out, n.value), list(st)


def prove_circuit(params, circ: Circuit, r, s, nthreads: int, inputs=None, aux=None, want_h: bool = False):
    """prove() on an oracle Circuit; params: bytes, numpy uint8 or ParamsBuf."""
    if isinstance(params, ParamsBuf):
        params = params.array
    return prove(params, circ.n_gates, circ.n_in, circ.n_aux, circ.rowptr, circ.col, circ.coef,
                 circ.inputs if inputs is None else inputs, circ.aux if aux is None else aux, r, s, nthreads, want_h)
