"""Host-side mirror of `fawkes_crypto::backend::bellman_groth16` over the C ABI.

Same names and argument meaning as the reference interface:
    setup      fawkes-crypto/src/backend/bellman_groth16/setup.rs:7-35
    prove      fawkes-crypto/src/backend/bellman_groth16/prover.rs:63-90
    verify     fawkes-crypto/src/backend/bellman_groth16/verifier.rs:75-81
    Parameters fawkes-crypto/src/backend/bellman_groth16/mod.rs:139-177  (write/read framing)
    Proof      prover.rs:13-60      VK  verifier.rs:12-73      G1Point/G2Point  group.rs:15-123
The circuit front-end (BuildCS/WitnessCS, circuit/r1cs/cs.rs) is out of scope: where the
reference runs the circuit closure to fill `values_input` / `values_aux`
(prover.rs:69-74), callers here pass those two vectors directly as uint64[n,4] arrays of
Montgomery limbs (the in-memory `Vec<Num<Fr>>`, cs.rs:99-102).

All arithmetic happens in libfawkes_b200.so (CUDA); this file only frames bytes.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import native as nv

FR_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_R256 = 1 << 256


def _to_mont(x: int, mod: int) -> int:
    return (x * _R256) % mod


def _from_mont(x: int, mod: int) -> int:
    return (x * pow(_R256, -1, mod)) % mod


def fr_raw(x: int) -> np.ndarray:
    """Canonical integer -> Num<Fr> in memory (uint64[4], Montgomery)."""
    return np.frombuffer(_to_mont(x % FR_MOD, FR_MOD).to_bytes(32, "little"), dtype=np.uint64).copy()


def fr_array(xs: Sequence[int]) -> np.ndarray:
    out = np.empty((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i] = fr_raw(x)
    return out


def fr_unraw(a) -> int:
    return _from_mont(int.from_bytes(np.asarray(a, dtype=np.uint64).tobytes(), "little"), FR_MOD)


class Context:
    """One CUDA device (fb_ctx).  One process per GPU."""

    def __init__(self, device: int = 0):
        self.handle = C.c_void_p()
        dev = (C.c_int * 1)(device)
        nv.check(nv.lib.fb_init(dev, 1, C.byref(self.handle)))
        self.device = device

    def close(self):
        if self.handle:
            nv.lib.fb_shutdown(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Circuit:
    """Parsed R1CS (fb_circuit): what WitnessCS::get_gate_iterator streams (cs.rs:248-250)."""

    def __init__(self, handle):
        self.handle = handle

    @classmethod
    def from_gates_blob(cls, blob: bytes, num_gates: int, n_in: int, n_aux: int, ctx: "Context" = None) -> "Circuit":
        """`Parameters.2` (brotli of the borsh gate stream).  With a Context the per-term work runs on the GPU
        (fb_circuit_from_gates_gpu; `ingest_ms` then holds the stage times); both build the same circuit."""
        h = C.c_void_p()
        if ctx is None:
            nv.check(nv.lib.fb_circuit_from_gates(nv.ptr(blob), len(blob), num_gates, n_in, n_aux, C.byref(h)))
            return cls(h)
        ms = (C.c_float * 6)()
        nv.check(nv.lib.fb_circuit_from_gates_gpu(ctx.handle, nv.ptr(blob), len(blob), num_gates, n_in, n_aux,
                                                  C.byref(h), ms))
        c = cls(h)
        c.ingest_ms = dict(zip(("frame", "upload", "kernels", "download", "brotli", "alloc"), [float(x) for x in ms]))
        return c

    @classmethod
    def from_raw_gates(cls, raw: bytes, num_gates: int, n_in: int, n_aux: int, ctx: "Context" = None) -> "Circuit":
        h = C.c_void_p()
        if ctx is None:
            nv.check(nv.lib.fb_circuit_from_raw_gates(nv.ptr(raw), len(raw), num_gates, n_in, n_aux, C.byref(h)))
            return cls(h)
        ms = (C.c_float * 6)()
        nv.check(nv.lib.fb_circuit_from_raw_gates_gpu(ctx.handle, nv.ptr(raw), len(raw), num_gates, n_in, n_aux,
                                                      C.byref(h), ms))
        c = cls(h)
        c.ingest_ms = dict(zip(("frame", "upload", "kernels", "download", "brotli", "alloc"), [float(x) for x in ms]))
        return c

    @classmethod
    def synthetic(cls, n_rows: int, seed: int) -> "Circuit":
        h = C.c_void_p()
        nv.check(nv.lib.fb_circuit_synth(n_rows, seed, C.byref(h)))
        return cls(h)

    def shape(self):
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        nv.check(nv.lib.fb_circuit_shape(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(n_in=a.value, n_aux=b.value, n_gates=c.value, nnz=d.value)

    def witness(self):
        """(inputs, aux) uint64[n,4] views for synthetic circuits (owned by the circuit)."""
        pi, pa = C.c_void_p(), C.c_void_p()
        nv.check(nv.lib.fb_circuit_witness(self.handle, C.byref(pi), C.byref(pa)))
        sh = self.shape()
        mk = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n, 4))
        return mk(pi, sh["n_in"]), mk(pa, sh["n_aux"])

    def close(self):
        if self.handle:
            nv.lib.fb_circuit_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ points ---
@dataclass
class G1Point:
    """(x, y) as Num<Fq> raw bytes; all-zero = infinity (group.rs:53-81)."""
    raw: bytes  # 64 B

    def canonical(self):
        v = [_from_mont(int.from_bytes(self.raw[i:i + 32], "little"), FQ_MOD) for i in (0, 32)]
        return tuple(v)

    def borsh(self) -> bytes:       # group.rs:15-31: two canonical 32 B LE numbers
        return b"".join(c.to_bytes(32, "little") for c in self.canonical())

    @classmethod
    def unborsh(cls, b: bytes) -> "G1Point":
        vals = [int.from_bytes(b[i:i + 32], "little") for i in (0, 32)]
        if any(v >= FQ_MOD for v in vals):
            raise ValueError("Wrong raw integer")
        return cls(b"".join(_to_mont(v, FQ_MOD).to_bytes(32, "little") for v in vals))


@dataclass
class G2Point:
    """((x.re, x.im), (y.re, y.im)) raw, 128 B (group.rs:83-123)."""
    raw: bytes

    def canonical(self):
        v = [_from_mont(int.from_bytes(self.raw[i:i + 32], "little"), FQ_MOD) for i in (0, 32, 64, 96)]
        return ((v[0], v[1]), (v[2], v[3]))

    def borsh(self) -> bytes:
        (a, b), (c, d) = self.canonical()
        return b"".join(x.to_bytes(32, "little") for x in (a, b, c, d))

    @classmethod
    def unborsh(cls, b: bytes) -> "G2Point":
        vals = [int.from_bytes(b[i:i + 32], "little") for i in (0, 32, 64, 96)]
        if any(v >= FQ_MOD for v in vals):
            raise ValueError("Wrong raw integer")
        return cls(b"".join(_to_mont(v, FQ_MOD).to_bytes(32, "little") for v in vals))


@dataclass
class Proof:
    a: G1Point
    b: G2Point
    c: G1Point

    def to_raw(self) -> bytes:
        return self.a.raw + self.b.raw + self.c.raw

    @classmethod
    def from_raw(cls, raw: bytes) -> "Proof":
        return cls(G1Point(raw[:64]), G2Point(raw[64:192]), G1Point(raw[192:256]))

    def serialize(self) -> bytes:               # prover.rs:38-45
        return self.a.borsh() + self.b.borsh() + self.c.borsh()

    @classmethod
    def deserialize(cls, b: bytes) -> "Proof":  # prover.rs:47-60
        return cls(G1Point.unborsh(b[:64]), G2Point.unborsh(b[64:192]), G1Point.unborsh(b[192:256]))


def _g1_from_be(b: bytes) -> G1Point:
    if b[0] & 0x40:
        return G1Point(bytes(64))
    x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:64], "big")
    return G1Point(_to_mont(x, FQ_MOD).to_bytes(32, "little") + _to_mont(y, FQ_MOD).to_bytes(32, "little"))


def _g2_from_be(b: bytes) -> G2Point:
    if b[0] & 0x40:
        return G2Point(bytes(128))
    x1, x0, y1, y0 = (int.from_bytes(b[32 * i:32 * i + 32], "big") for i in range(4))
    return G2Point(b"".join(_to_mont(v, FQ_MOD).to_bytes(32, "little") for v in (x0, x1, y0, y1)))


@dataclass
class VK:
    alpha: G1Point
    beta: G2Point
    gamma: G2Point
    delta: G2Point
    ic: List[G1Point]

    def to_raw(self) -> bytes:
        return self.alpha.raw + self.beta.raw + self.gamma.raw + self.delta.raw + b"".join(p.raw for p in self.ic)

    def serialize(self) -> bytes:               # verifier.rs:45-54
        return (self.alpha.borsh() + self.beta.borsh() + self.gamma.borsh() + self.delta.borsh() +
                struct.pack("<I", len(self.ic)) + b"".join(p.borsh() for p in self.ic))

    @classmethod
    def deserialize(cls, b: bytes) -> "VK":     # verifier.rs:56-73
        alpha = G1Point.unborsh(b[:64])
        beta, gamma, delta = (G2Point.unborsh(b[64 + 128 * i:192 + 128 * i]) for i in range(3))
        (n,) = struct.unpack_from("<I", b, 448)
        ic = [G1Point.unborsh(b[452 + 64 * i:516 + 64 * i]) for i in range(n)]
        return cls(alpha, beta, gamma, delta, ic)


# -------------------------------------------------------------- Parameters ---
def _bitvec_to_bytes(bits) -> bytes:            # bit_vec::BitVec::to_bytes (MSB first)
    out = bytearray((len(bits) + 7) // 8)
    for i, b in enumerate(bits):
        if b:
            out[i // 8] |= 0x80 >> (i % 8)
    return bytes(out)


class Parameters:
    """`Parameters(bellman params, num_gates, gates blob, const tracker)` (mod.rs:139)."""

    def __init__(self, bellman_bytes: bytes, num_gates: int, gates_blob: bytes, const_tracker=(),
                 circuit: Optional[Circuit] = None):
        self.bellman_bytes = bellman_bytes      # bytes, or a uint8 numpy array for multi-GiB keys
        self.num_gates = num_gates
        self.gates_blob = gates_blob
        self.const_tracker = list(const_tracker)
        self._circuit = circuit
        self._pk = None
        self._pk_key = None     # (ctx, shard, nshards, flags) the resident key was loaded with
        self._checked = True
        self._disallow_inf = False

    # -- (de)serialisation, mod.rs:150-175 ------------------------------------
    def write(self) -> bytes:
        bv = _bitvec_to_bytes(self.const_tracker)
        return (struct.pack("<I", self.num_gates) + struct.pack("<I", len(self.gates_blob)) + self.gates_blob +
                struct.pack("<I", len(self.const_tracker)) + struct.pack("<I", len(bv)) + bv +
                bytes(self.bellman_bytes))

    @classmethod
    def read(cls, data: bytes, disallow_points_at_infinity: bool = False, checked: bool = True) -> "Parameters":
        try:
            (num_gates, blen) = struct.unpack_from("<II", data, 0)
            blob = data[8:8 + blen]
            if len(blob) != blen:
                raise struct.error("short")
            pos = 8 + blen
            nbits, nbytes = struct.unpack_from("<II", data, pos)
            pos += 8
        except struct.error as e:
            raise IOError("unexpected end of Parameters") from e
        if nbits > nbytes * 8:
            raise IOError("inconsistent bitvec length")
        bv = data[pos:pos + nbytes]
        if len(bv) != nbytes:
            raise IOError("unexpected end of Parameters")
        pos += nbytes
        bits = [bool(bv[i // 8] & (0x80 >> (i % 8))) for i in range(nbits)]
        p = cls(data[pos:], num_gates, blob, bits)
        # the two flags of bellman's Parameters::read are applied where the points are decoded: at key load
        p._checked = bool(checked)
        p._disallow_inf = bool(disallow_points_at_infinity)
        return p

    # -- views -----------------------------------------------------------------
    def _sections(self):
        b = self.bellman_bytes
        (n_ic,) = struct.unpack_from(">I", b, 576)
        pos = 580 + 64 * n_ic
        out = {"n_ic": n_ic}
        for name, sz in (("h", 64), ("l", 64), ("a", 64), ("b_g1", 64), ("b_g2", 128)):
            (n,) = struct.unpack_from(">I", b, pos)
            out[name] = (pos + 4, n)
            pos += 4 + n * sz
        return out

    @property
    def n_in(self) -> int:
        return self._sections()["n_ic"]

    @property
    def n_aux(self) -> int:
        return self._sections()["l"][1]

    def get_vk(self) -> VK:                    # mod.rs:142-144
        n_ic = self._sections()["n_ic"]
        b = bytes(self.bellman_bytes[:580 + 64 * n_ic])
        return VK(_g1_from_be(b[0:64]), _g2_from_be(b[128:256]), _g2_from_be(b[256:384]),
                  _g2_from_be(b[448:576]), [_g1_from_be(b[580 + 64 * i:644 + 64 * i]) for i in range(n_ic)])

    def circuit(self, ctx: "Context" = None) -> Circuit:
        if self._circuit is None:
            self._circuit = Circuit.from_gates_blob(self.gates_blob, self.num_gates, self.n_in, self.n_aux, ctx)
        return self._circuit

    def load(self, ctx: Context, checked: Optional[bool] = None, shard: int = 0, nshards: int = 1):
        """Upload the proving key to HBM (fb_pk_load_shard).  One resident key per Parameters: asking for it
        again with another context, shard or flags is an error (unload() first) -- a key's arrays, streams and
        events belong to the context it was loaded on."""
        if checked is None and self._pk is not None and self._pk_key[:3] == (id(ctx), shard, nshards):
            return self._pk     # whatever flags the resident key was validated with
        checked = self._checked if checked is None else bool(checked)
        flags = (nv.FB_LOAD_CHECKED if checked else 0) | (nv.FB_LOAD_NO_INFINITY if self._disallow_inf else 0)
        key = (id(ctx), shard, nshards, flags)
        if self._pk is not None and self._pk_key != key:
            raise ValueError("Parameters.load: a key is already resident with another context / shard / flags; "
                             "call unload() first")
        if self._pk is None:
            h = C.c_void_p()
            nv.check(nv.lib.fb_pk_load_shard(ctx.handle, nv.ptr(self.bellman_bytes), len(self.bellman_bytes),
                                             self.circuit(ctx).handle, flags, shard, nshards, C.byref(h)))
            self._pk, self._pk_key, self._pk_ctx = h, key, ctx
        return self._pk

    def info(self) -> dict:
        inf = nv.PkInfo()
        nv.check(nv.lib.fb_pk_get_info(self._pk, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in nv.PkInfo._fields_}

    def timings(self) -> dict:
        ms = (C.c_float * 6)()
        nv.check(nv.lib.fb_prove_timings(self._pk, ms))
        return dict(zip(("h2d", "r1cs", "h_ntt", "msm", "host", "total"), [float(x) for x in ms]))

    def unload(self):
        if self._pk is not None:
            nv.lib.fb_pk_free(self._pk)
            self._pk = None
            self._pk_key = None

    def __del__(self):
        try:
            self.unload()
        except Exception:
            pass


# ------------------------------------------------------------ entry points ---
def _sample_fr() -> int:
    """One uniformly random Fr, drawn the way the reference draws r and s (prover.rs:78-80): `OsRng::next_u32` is 4
    bytes of OS randomness read BIG-endian (osrng.rs:12-18), `next_u64` is rand 0.4's default (two next_u32, high
    word first), ff_ce's `Rand for Fr` fills the 4 limbs in order, clears the top REPR_SHAVE_BITS = 2 bits of the
    top limb, rejects values >= r and takes the limbs AS the Montgomery representation.  Returned as the canonical
    integer of that Montgomery value (rust/fawkes-b200/src/osrng.rs is the same routine for the Rust shim)."""
    while True:
        limbs = []
        for _ in range(4):
            hi = int.from_bytes(os.urandom(4), "big")
            lo = int.from_bytes(os.urandom(4), "big")
            limbs.append((hi << 32) | lo)
        limbs[3] &= (1 << 62) - 1
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < FR_MOD:
            return _from_mont(v, FR_MOD)


def setup(circuit: Circuit, ctx: Context, trapdoor: Optional[Sequence[int]] = None, num_gates: Optional[int] = None,
          gates_blob: bytes = b"", const_tracker=(), shard: int = 0, nshards: int = 1) -> Parameters:
    """`setup(circuit)` of setup.rs:7-35 for an already-built R1CS.  trapdoor = (alpha, beta,
    gamma, delta, tau) canonical integers; sampled from the OS when omitted.  nshards > 1: only the query points
    of this rank's shard are generated (fb_setup_shard; every rank must pass the same trapdoor)."""
    td = list(trapdoor) if trapdoor is not None else [_sample_fr() for _ in range(5)]
    tda = fr_array(td)
    out, n = C.c_void_p(), C.c_size_t()
    if nshards == 1:
        nv.check(nv.lib.fb_setup(ctx.handle, circuit.handle, nv.ptr(tda), C.byref(out), C.byref(n)))
    else:
        nv.check(nv.lib.fb_setup_shard(ctx.handle, circuit.handle, nv.ptr(tda), shard, nshards, C.byref(out), C.byref(n)))
    try:
        # ctypes.string_at takes a C int length: copy by hand so keys beyond 2 GiB survive
        data = np.empty(n.value, dtype=np.uint8)
        C.memmove(data.ctypes.data, out, n.value)
    finally:
        nv.lib.fb_free(out)
    ng = circuit.shape()["n_gates"] if num_gates is None else num_gates
    return Parameters(data, ng, gates_blob, const_tracker, circuit=circuit)


def prove_with_rs(params: Parameters, values_input: np.ndarray, values_aux: np.ndarray, r: int, s: int,
                  ctx: Context, return_h: bool = False):
    """create_proof(circuit, params, r, s): deterministic blinding.  Returns
    (public inputs [without ONE], Proof) like prover.rs:84-89."""
    pk = params.load(ctx)
    vi = np.ascontiguousarray(values_input, dtype=np.uint64)
    va = np.ascontiguousarray(values_aux, dtype=np.uint64)
    ra, sa = fr_raw(r), fr_raw(s)
    out = np.zeros(256, dtype=np.uint8)
    h = None
    if return_h:
        m = 1 << params.info()["log_m"]
        h = np.zeros((m - 1, 4), dtype=np.uint64)
    nv.check(nv.lib.fb_prove(ctx.handle, pk, nv.ptr(vi), vi.shape[0], nv.ptr(va), va.shape[0],
                             nv.ptr(ra), nv.ptr(sa), nv.ptr(out), nv.ptr(h) if h is not None else None))
    proof = Proof.from_raw(out.tobytes())
    inputs = vi[1:].copy()
    return (inputs, proof, h) if return_h else (inputs, proof)


def prove_batch(params: Parameters, witnesses, rs: Sequence[int], ss: Sequence[int], ctx: Context):
    """Many proofs of ONE circuit on one resident key (BASELINE configs[1]); the reference proves them one
    `prove()` call at a time (prover.rs:63-90).  witnesses: sequence of (values_input, values_aux) pairs.
    Returns [(public inputs, Proof), ...] in order; proof i is byte-identical to
    prove_with_rs(params, *witnesses[i], rs[i], ss[i])."""
    count = len(witnesses)
    if not (count == len(rs) == len(ss)):
        raise ValueError("prove_batch: witnesses, rs and ss differ in length")
    if count == 0:
        return []
    pk = params.load(ctx)
    vis = [np.ascontiguousarray(w[0], dtype=np.uint64) for w in witnesses]
    vas = [np.ascontiguousarray(w[1], dtype=np.uint64) for w in witnesses]
    n_in, n_aux = vis[0].shape[0], vas[0].shape[0]
    if any(v.shape != (n_in, 4) for v in vis) or any(v.shape != (n_aux, 4) for v in vas):
        raise ValueError("prove_batch: every witness must have the circuit's shape")
    ins = (C.c_void_p * count)(*[v.ctypes.data for v in vis])
    axs = (C.c_void_p * count)(*[v.ctypes.data for v in vas])
    ra, sa = fr_array(list(rs)), fr_array(list(ss))
    out = np.zeros((count, 256), dtype=np.uint8)
    nv.check(nv.lib.fb_prove_batch(ctx.handle, pk, count, ins, n_in, axs, n_aux, nv.ptr(ra), nv.ptr(sa), nv.ptr(out)))
    return [(vis[i][1:].copy(), Proof.from_raw(out[i].tobytes())) for i in range(count)]


class ProveStream:
    """Streaming proves on one resident key (fb_stream_*): `submit` copies the witness and returns a ticket,
    so the caller builds witness k+1 (the circuit closure re-run of prover.rs:69-76) while proof k is on the
    GPU; `wait(ticket)` returns (public inputs, Proof), byte-identical to prove_with_rs on the same arguments.
    While the stream is open the key is used through it only."""

    def __init__(self, params: Parameters, ctx: Context, depth: int = 0):
        self._pk = params.load(ctx)
        self._params = params
        self._h = C.c_void_p()
        self._inputs = {}
        nv.check(nv.lib.fb_stream_open(ctx.handle, self._pk, depth, C.byref(self._h)))

    def submit(self, values_input: np.ndarray, values_aux: np.ndarray, r: int, s: int) -> int:
        vi = np.ascontiguousarray(values_input, dtype=np.uint64)
        va = np.ascontiguousarray(values_aux, dtype=np.uint64)
        ra, sa = fr_raw(r), fr_raw(s)
        t = C.c_uint64()
        nv.check(nv.lib.fb_stream_submit(self._h, nv.ptr(vi), vi.shape[0], nv.ptr(va), va.shape[0], nv.ptr(ra), nv.ptr(sa),
                                         C.byref(t)))
        self._inputs[t.value] = vi[1:].copy()
        return t.value

    def wait(self, ticket: int):
        out = np.zeros(256, dtype=np.uint8)
        nv.check(nv.lib.fb_stream_wait(self._h, ticket, nv.ptr(out)))
        return self._inputs.pop(ticket), Proof.from_raw(out.tobytes())

    def close(self):
        if self._h:
            nv.lib.fb_stream_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prove(params: Parameters, values_input: np.ndarray, values_aux: np.ndarray, ctx: Context):
    """`prove` of prover.rs:63-90: r, s from the OS RNG."""
    return prove_with_rs(params, values_input, values_aux, _sample_fr(), _sample_fr(), ctx)


def verify(vk: VK, proof: Proof, inputs: np.ndarray) -> bool:
    """`verify` of verifier.rs:75-81.  inputs: uint64[n,4] Montgomery, without ONE.
    Raises on a length mismatch, and on limbs that are not reduced field elements or points that are not on
    their curve in the key, the proof or the inputs (the reference panics there through `.unwrap()`:
    verifier.rs:80, group.rs:53-65,87-105, mod.rs:105-120)."""
    raw = vk.to_raw()
    ins = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1, 4)
    ok = C.c_int()
    pr = proof.to_raw()
    nv.check(nv.lib.fb_verify(nv.ptr(raw), len(vk.ic), nv.ptr(pr), nv.ptr(ins) if len(ins) else None,
                              ins.shape[0], C.byref(ok)))
    return bool(ok.value)
