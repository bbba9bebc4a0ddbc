"""Front-end restatement for configs[0]: the Poseidon Merkle-proof circuit (depth 32, 7,328 constraints).

TEST INFRASTRUCTURE ONLY: builds the R1CS and the witness that the
reference's circuit DSL would stream into the prover for `tests/bellman_groth16.rs:19-47`, so that the
hot path can be exercised on the real cfg-1 shape (rows with up to ~55 terms, 7,364 rows, m = 2^13)
instead of synthetic 3-term rows.  Nothing under fawkes-crypto_b200/ imports this module.

What is restated, with the reference lines it follows (paths relative to /root/reference):
  * LC = sorted list of (coeff, Index), Input < Aux, zero coefficients dropped
    fawkes-crypto/src/circuit/r1cs/lc.rs:42-135,138-163; a constant is a term on Input(0)
  * CNum add/sub/mul-by-constant are free, CNum * CNum allocates one aux and one gate unless a side is
    constant                                   circuit/r1cs/num.rs:161-171,228-272
  * alloc / inputize / enforce order           circuit/r1cs/cs.rs:309-329, backend/bellman_groth16/prover.rs:69-72
  * CBool::alloc = alloc + assert_bit b (b - 1) = 0     circuit/r1cs/bool.rs:68-71, num.rs:81-83
  * Poseidon permutation (ark, x^5 as 3 products, MDS mix), c_poseidon, merkle root
    circuit/poseidon.rs:16-95, native/poseidon.rs:24-125
  * Poseidon parameters: seed "fawkes_poseidon(t=..,f=..,p=..,salt=..)" -> keccak256 -> ChaCha20Rng,
    draws c[f+p][t], x[t], y[t], m[i][j] = 1 / (x_i + y_j); a draw fills 4 u64 limbs, masks the top limb
    to 62 bits, rejects >= r and interprets the limbs AS MONTGOMERY FORM
    native/poseidon.rs:24-48, seedbox/src/lib.rs:9-38, ff-uint/src/num/mod.rs:286-303
Parity status: the reference holds no known-answer vectors for the Poseidon constants or hashes (the
ChaCha20Rng / keccak crates are un-vendored too), so the constants are "restated, unpinned"; what IS pinned
is the circuit shape: 228 gates per hash, 7,328 for the 32 levels (README.md:52) -- tests/test_frontend.py.
"""
from __future__ import annotations

import struct

from .bn254 import R
from .groth16 import INPUT, AUX

MASK64 = (1 << 64) - 1
R_MONT_INV = pow(1 << 256, -1, R)


# ------------------------------------------------------------------ keccak256 ---
_KECCAK_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
    0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
    0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
    0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_KECCAK_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol64(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & MASK64 if n else x


def _keccak_f(a):
    for rc in _KECCAK_RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol64(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol64(a[x][y], _KECCAK_ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    """Original Keccak-256 (pad 0x01 .. 0x80, rate 136), as the `sha3::Keccak256` the seedbox uses."""
    rate = 136
    msg = bytearray(data) + b"\x01"
    msg += b"\x00" * ((-len(msg)) % rate)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= struct.unpack_from("<Q", msg, off + 8 * i)[0]
        a = _keccak_f(a)
    return b"".join(struct.pack("<Q", a[i % 5][i // 5]) for i in range(4))


# ------------------------------------------------------------------ ChaCha20Rng ---
def chacha20_block(key_words, w12, w13, w14, w15):
    """One ChaCha20 block (20 rounds) -> 16 output words."""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [w12, w13, w14, w15]
    x = st[:]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & 0xFFFFFFFF
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & 0xFFFFFFFF

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class SeedboxChaCha20:
    """seedbox/src/lib.rs:20-38: ChaCha20Rng::from_seed(keccak256(salt)); 64-bit block counter from 0,
    stream id 0; next_u64 = two consecutive output words, low word first (rand_core BlockRng)."""

    def __init__(self, salt: bytes):
        self.key = struct.unpack("<8I", keccak256(salt))
        self.counter = 0
        self.buf = []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter & 0xFFFFFFFF, self.counter >> 32, 0, 0)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)

    def gen_fr(self) -> int:
        """ff-uint/src/num/mod.rs:286-303: limbs -> mask -> reject -> Montgomery limbs; returns the
        canonical value."""
        while True:
            l = [self.next_u64() for _ in range(4)]
            l[3] &= MASK64 >> 2          # REPR_SHAVE_BITS = 2 for a 254-bit modulus
            v = l[0] | (l[1] << 64) | (l[2] << 128) | (l[3] << 192)
            if v < R:
                return v * R_MONT_INV % R


class PoseidonParams:
    def __init__(self, t, f, p, salt=""):
        sb = SeedboxChaCha20(f"fawkes_poseidon(t={t},f={f},p={p},salt={salt})".encode())
        self.t, self.f, self.p = t, f, p
        self.c = [[sb.gen_fr() for _ in range(t)] for _ in range(f + p)]
        x = [sb.gen_fr() for _ in range(t)]
        y = [sb.gen_fr() for _ in range(t)]
        self.m = [[pow((x[i] + y[j]) % R, -1, R) for j in range(t)] for i in range(t)]


def poseidon(inputs, params: PoseidonParams) -> int:
    """native/poseidon.rs:50-98"""
    assert 0 < len(inputs) < params.t
    st = list(inputs) + [0] * (params.t - len(inputs))
    half = params.f >> 1
    for i in range(params.f + params.p):
        st = [(s + c) % R for s, c in zip(st, params.c[i])]
        if i < half or i >= half + params.p:
            st = [pow(s, 5, R) for s in st]
        else:
            st[0] = pow(st[0], 5, R)
        st = [sum(params.m[a][b] * st[b] for b in range(params.t)) % R for a in range(params.t)]
    return st[0]


def poseidon_merkle_proof_root(leaf, sibling, path, params) -> int:
    """native/poseidon.rs:115-125"""
    root = leaf
    for p, s in zip(path, sibling):
        root = poseidon([s, root] if p else [root, s], params)
    return root


# ------------------------------------------------------------------ circuit DSL ---
class BuildCS:
    """Gates + witness in one pass (the reference runs the closure twice, once per CS kind;
    the gate stream and the variable numbering are the same)."""

    def __init__(self):
        self.gates = []
        self.inputs = [1]      # values_input[0] = ONE (cs.rs:111)
        self.aux = []

    def alloc(self, value) -> "CNum":
        self.aux.append(value % R)
        return CNum(self, {(AUX, len(self.aux) - 1): 1}, value % R)

    def const(self, value) -> "CNum":
        value %= R
        return CNum(self, {(INPUT, 0): value} if value else {}, value)

    def enforce(self, a: "CNum", b: "CNum", c: "CNum"):
        assert a.value * b.value % R == c.value, "unsatisfied gate"
        self.gates.append((a.terms(), b.terms(), c.terms()))

    def inputize(self, n: "CNum"):
        """cs.rs:309-318"""
        self.inputs.append(n.value)
        self.gates.append((n.terms(), [(1, (INPUT, 0))], [(1, (INPUT, len(self.inputs) - 1))]))


class CNum:
    def __init__(self, cs, lc, value):
        self.cs, self.lc, self.value = cs, lc, value

    def terms(self):
        return [(c, k) for k, c in sorted(self.lc.items())]     # Input < Aux, ascending (lc.rs:138-163)

    def as_const(self):
        if not self.lc:
            return 0
        if len(self.lc) == 1 and (INPUT, 0) in self.lc:
            return self.lc[(INPUT, 0)]
        return None

    def _lin(self, other, sign):
        lc = dict(self.lc)
        for k, c in other.lc.items():
            v = (lc.get(k, 0) + sign * c) % R
            if v:
                lc[k] = v
            else:
                lc.pop(k, None)
        return CNum(self.cs, lc, (self.value + sign * other.value) % R)

    def __add__(self, other):
        return self._lin(other if isinstance(other, CNum) else self.cs.const(other), 1)

    def __sub__(self, other):
        return self._lin(other if isinstance(other, CNum) else self.cs.const(other), -1)

    def scale(self, k):
        k %= R
        if k == 0:
            return self.cs.const(0)
        return CNum(self.cs, {i: c * k % R for i, c in self.lc.items()}, self.value * k % R)

    def __mul__(self, other):
        """num.rs:253-272"""
        if not isinstance(other, CNum):
            return self.scale(other)
        a, b = self.as_const(), other.as_const()
        if a is not None:
            return other.scale(a)
        if b is not None:
            return self.scale(b)
        out = self.cs.alloc(self.value * other.value % R)
        self.cs.enforce(self, other, out)
        return out

    def assert_bit(self):
        self.cs.enforce(self, self - 1, self.cs.const(0))       # num.rs:81-83

    def assert_eq(self, other):
        self.cs.enforce(self, self.cs.const(1), other)          # num.rs:173-175

    def switch(self, bit: "CNum", if_else: "CNum") -> "CNum":
        """num.rs:161-171: bit ? self : if_else"""
        return if_else + (self - if_else) * bit


def c_poseidon(inputs, params: PoseidonParams) -> CNum:
    """circuit/poseidon.rs:16-66"""
    cs = inputs[0].cs
    st = list(inputs) + [cs.const(0)] * (params.t - len(inputs))
    half = params.f >> 1

    def sigma(a):
        sq = a * a
        return (sq * sq) * a

    for i in range(params.f + params.p):
        st = [s + c for s, c in zip(st, params.c[i])]
        if i < half or i >= half + params.p:
            st = [sigma(s) for s in st]
        else:
            st[0] = sigma(st[0])
        new = []
        for a in range(params.t):
            acc = cs.const(0)
            for b in range(params.t):
                acc = acc + st[b].scale(params.m[a][b])
            new.append(acc)
        st = new
    return st[0]


def c_poseidon_merkle_proof_root(leaf, sibling, path, params) -> CNum:
    """circuit/poseidon.rs:83-95"""
    root = leaf
    for p, s in zip(path, sibling):
        first = s.switch(p, root)
        second = root + s - first
        root = c_poseidon([first, second], params)
    return root


def merkle_circuit(leaf: int, sibling, path, params: PoseidonParams = None):
    """The whole of tests/bellman_groth16.rs:19-47 as the prover sees it: Pub = root (alloc + inputize,
    prover.rs:70-71), Sec = (leaf, CMerkleProof{sibling, path}) allocated field by field, then the circuit.
    Returns (gates, values_input, values_aux)."""
    params = params or PoseidonParams(3, 8, 53)
    root = poseidon_merkle_proof_root(leaf, sibling, path, params)
    cs = BuildCS()
    c_root = cs.alloc(root)
    cs.inputize(c_root)
    c_leaf = cs.alloc(leaf)
    c_sib = [cs.alloc(s) for s in sibling]
    c_path = []
    for p in path:
        b = cs.alloc(1 if p else 0)
        b.assert_bit()
        c_path.append(b)
    res = c_poseidon_merkle_proof_root(c_leaf, c_sib, c_path, params)
    res.assert_eq(c_root)
    return cs.gates, cs.inputs, cs.aux
