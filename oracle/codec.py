"""Wire formats of the groth16 backend boundary -- Python restatement.

TEST INFRASTRUCTURE ONLY (see oracle/bn254.py header for the import rule).

Follows:
  * Num<Fp> borsh: 32 B canonical LE          ff-uint_derive/src/lib.rs:687-702
  * Index borsh: u8 tag + u32 LE              fawkes-crypto/src/circuit/r1cs/lc.rs:144-149
  * Gate borsh: 3 x (u32 LE n + n x 37 B)     circuit/r1cs/cs.rs:20-26,186-213
  * gates blob: brotli(4096, q9, lgwin 22)    backend/bellman_groth16/setup.rs:25-32
  * Parameters framing                        backend/bellman_groth16/mod.rs:150-175
  * raw points (FFI / Proof in memory)        backend/bellman_groth16/group.rs:53-123
  * Proof / VK borsh                          prover.rs:38-60, verifier.rs:45-73
  * bellman Parameters body: big-endian uncompressed points, u32 BE lengths
    [restated from bellman_ce 0.3.5 / pairing_ce 0.18.1, absent from /root/reference]
"""
from __future__ import annotations

import ctypes
import ctypes.util
import struct

from . import bn254 as bn
from .bn254 import P, R
from .groth16 import Params, VerifyingKey, Proof, INPUT, AUX

# ---------------------------------------------------------------- brotli ---
_enc = _dec = None


def _load_brotli():
    global _enc, _dec
    if _enc is None:
        _enc = ctypes.CDLL("libbrotlienc.so.1")
        _dec = ctypes.CDLL("libbrotlidec.so.1")
        _enc.BrotliEncoderCompress.restype = ctypes.c_int
        _enc.BrotliEncoderCompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                               ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
        _enc.BrotliEncoderMaxCompressedSize.restype = ctypes.c_size_t
        _enc.BrotliEncoderMaxCompressedSize.argtypes = [ctypes.c_size_t]
        _dec.BrotliDecoderDecompress.restype = ctypes.c_int
        _dec.BrotliDecoderDecompress.argtypes = [ctypes.c_size_t, ctypes.c_char_p,
                                                 ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]


def brotli_compress(data: bytes, quality=9, lgwin=22) -> bytes:
    _load_brotli()
    cap = _enc.BrotliEncoderMaxCompressedSize(len(data)) or (len(data) + 1024)
    out = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(cap)
    ok = _enc.BrotliEncoderCompress(quality, lgwin, 0, len(data), data, ctypes.byref(n), out)
    assert ok == 1
    return out.raw[:n.value]


def brotli_decompress(data: bytes, max_out: int) -> bytes:
    _load_brotli()
    out = ctypes.create_string_buffer(max_out)
    n = ctypes.c_size_t(max_out)
    ok = _dec.BrotliDecoderDecompress(len(data), data, ctypes.byref(n), out)
    assert ok == 1, "brotli decode failed"
    return out.raw[:n.value]


# ------------------------------------------------------------ field bytes ---
def num_borsh(x: int) -> bytes:
    return x.to_bytes(32, "little")


def num_unborsh(b: bytes, mod=R) -> int:
    v = int.from_bytes(b, "little")
    if v >= mod:
        raise ValueError("Wrong raw integer")
    return v


def fr_raw(x: int) -> bytes:
    """In-memory Num<Fr>: 4xu64 LE Montgomery (ff-uint/src/num/mod.rs:21-23)."""
    return bn.fr_to_mont(x).to_bytes(32, "little")


def fr_unraw(b: bytes) -> int:
    return bn.fr_from_mont(int.from_bytes(b, "little"))


def fq_raw(x): return bn.fq_to_mont(x).to_bytes(32, "little")
def fq_unraw(b): return bn.fq_from_mont(int.from_bytes(b, "little"))


# ------------------------------------------------------------------ gates ---
def gate_borsh(gate) -> bytes:
    out = bytearray()
    for lc in gate:
        out += struct.pack("<I", len(lc))
        for coeff, (tag, idx) in lc:
            out += num_borsh(coeff) + struct.pack("<BI", tag, idx)
    return bytes(out)


def gates_blob(gates) -> bytes:
    return brotli_compress(b"".join(gate_borsh(g) for g in gates))


def parse_gates(raw: bytes):
    """GateStreamedIterator semantics: stop silently at the first short read
    (circuit/r1cs/cs.rs:215-223)."""
    gates, pos = [], 0
    while True:
        gate, ok = [], True
        for _ in range(3):
            if pos + 4 > len(raw):
                ok = False
                break
            (n,) = struct.unpack_from("<I", raw, pos)
            pos += 4
            if pos + 37 * n > len(raw):
                ok = False
                break
            lc = []
            for _k in range(n):
                c = num_unborsh(raw[pos:pos + 32])
                tag, idx = struct.unpack_from("<BI", raw, pos + 32)
                if tag > 1:
                    raise ValueError("enum elements overflow")
                lc.append((c, (tag, idx)))
                pos += 37
            gate.append(lc)
        if not ok:
            return gates
        gates.append(tuple(gate))


# --------------------------------------------------- bellman point encoding ---
def g1_uncompressed(pt) -> bytes:
    if pt is None:
        return bytes([0x40]) + bytes(63)
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def g2_uncompressed(pt) -> bytes:
    if pt is None:
        return bytes([0x40]) + bytes(127)
    (x0, x1), (y0, y1) = pt
    return b"".join(v.to_bytes(32, "big") for v in (x1, x0, y1, y0))


def _fq_be(b):
    v = int.from_bytes(b, "big")
    if v >= P:
        raise ValueError("coordinate not in field")
    return v


def g1_from_uncompressed(b: bytes):
    if b[0] & 0x40:
        return None
    return (_fq_be(b[:32]), _fq_be(b[32:64]))


def g2_from_uncompressed(b: bytes):
    if b[0] & 0x40:
        return None
    x1, x0, y1, y0 = (_fq_be(b[32 * i:32 * i + 32]) for i in range(4))
    return ((x0, x1), (y0, y1))


def bellman_params_bytes(p: Params) -> bytes:
    vk = p.vk
    out = bytearray()
    out += g1_uncompressed(vk.alpha_g1) + g1_uncompressed(vk.beta_g1)
    out += g2_uncompressed(vk.beta_g2) + g2_uncompressed(vk.gamma_g2)
    out += g1_uncompressed(vk.delta_g1) + g2_uncompressed(vk.delta_g2)
    out += struct.pack(">I", len(vk.ic)) + b"".join(g1_uncompressed(x) for x in vk.ic)
    for name, enc in (("h", g1_uncompressed), ("l", g1_uncompressed), ("a", g1_uncompressed),
                      ("b_g1", g1_uncompressed), ("b_g2", g2_uncompressed)):
        pts = getattr(p, name)
        out += struct.pack(">I", len(pts)) + b"".join(enc(x) for x in pts)
    return bytes(out)


def bellman_params_parse(b: bytes) -> Params:
    pos = 0

    def g1():
        nonlocal pos
        pos += 64
        return g1_from_uncompressed(b[pos - 64:pos])

    def g2():
        nonlocal pos
        pos += 128
        return g2_from_uncompressed(b[pos - 128:pos])

    def vec(f):
        nonlocal pos
        (n,) = struct.unpack_from(">I", b, pos)
        pos += 4
        return [f() for _ in range(n)]

    a1, b1, b2, g2_, d1, d2 = g1(), g1(), g2(), g2(), g1(), g2()
    ic = vec(g1)
    vk = VerifyingKey(a1, b1, b2, g2_, d1, d2, ic)
    h, l, a, bg1, bg2 = vec(g1), vec(g1), vec(g1), vec(g1), vec(g2)
    assert pos == len(b)
    return Params(vk, h, l, a, bg1, bg2)


# ------------------------------------------- fawkes Parameters framing ---
def bitvec_bytes(bits) -> bytes:
    """bit_vec::BitVec::to_bytes: MSB-first within each byte."""
    out = bytearray((len(bits) + 7) // 8)
    for i, b in enumerate(bits):
        if b:
            out[i // 8] |= 0x80 >> (i % 8)
    return bytes(out)


def fawkes_params_bytes(p: Params, num_gates: int, blob: bytes, const_tracker) -> bytes:
    bv = bitvec_bytes(const_tracker)
    return (struct.pack("<I", num_gates) + struct.pack("<I", len(blob)) + blob +
            struct.pack("<I", len(const_tracker)) + struct.pack("<I", len(bv)) + bv +
            bellman_params_bytes(p))


def fawkes_params_parse(b: bytes):
    (num_gates,) = struct.unpack_from("<I", b, 0)
    (blen,) = struct.unpack_from("<I", b, 4)
    blob = b[8:8 + blen]
    pos = 8 + blen
    nbits, nbytes = struct.unpack_from("<II", b, pos)
    pos += 8
    if nbits > nbytes * 8:
        raise ValueError("inconsistent bitvec length")
    bv = b[pos:pos + nbytes]
    pos += nbytes
    bits = [bool(bv[i // 8] & (0x80 >> (i % 8))) for i in range(nbits)]
    return bellman_params_parse(b[pos:]), num_gates, blob, bits


# ------------------------------------------------------- raw / borsh proof ---
def g1_raw(pt) -> bytes:
    """G1Point in memory: x,y raw LE Montgomery; infinity = zeros (group.rs:53-81)."""
    if pt is None:
        return bytes(64)
    return fq_raw(pt[0]) + fq_raw(pt[1])


def g2_raw(pt) -> bytes:
    """G2Point: x.c0 x.c1 y.c0 y.c1 (group.rs:83-123)."""
    if pt is None:
        return bytes(128)
    return fq_raw(pt[0][0]) + fq_raw(pt[0][1]) + fq_raw(pt[1][0]) + fq_raw(pt[1][1])


def g1_unraw(b):
    if b == bytes(64):
        return None
    return (fq_unraw(b[:32]), fq_unraw(b[32:]))


def g2_unraw(b):
    if b == bytes(128):
        return None
    v = [fq_unraw(b[32 * i:32 * i + 32]) for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


def proof_raw(pr: Proof) -> bytes:
    return g1_raw(pr.a) + g2_raw(pr.b) + g1_raw(pr.c)


def proof_unraw(b: bytes) -> Proof:
    return Proof(g1_unraw(b[:64]), g2_unraw(b[64:192]), g1_unraw(b[192:256]))


def proof_borsh(pr: Proof) -> bytes:
    """Canonical 32 B LE per coordinate (prover.rs:38-60, group.rs:15-51)."""
    def c(v): return (v or 0).to_bytes(32, "little")
    a = pr.a or (0, 0)
    b = pr.b or ((0, 0), (0, 0))
    cc = pr.c or (0, 0)
    return c(a[0]) + c(a[1]) + c(b[0][0]) + c(b[0][1]) + c(b[1][0]) + c(b[1][1]) + c(cc[0]) + c(cc[1])
