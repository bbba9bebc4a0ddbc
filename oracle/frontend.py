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


# ====================================================================================================
# configs[1]: the EdDSA-Poseidon signature circuit (`c_eddsaposeidon_verify`, circuit/eddsaposeidon.rs:16-47)
#
# Restated below, with the reference lines each piece follows:
#   * JubJub-on-BN254 parameters d, A, B, u and the generator from seed "edwards_g"
#     engines/bn256/mod.rs:28-72, constants.rs:1-2, native/ecc.rs:104-135 (from_scalar_raw)
#   * native twisted-Edwards / Montgomery arithmetic, subgroup_decompress      native/ecc.rs:54-231,322-353
#   * EdDSA-Poseidon sign / verify (Blake2s nonce, personalisation "__fawkes")   native/eddsaposeidon.rs:14-79
#   * CNum::div_unchecked / is_zero / assert_const, CBool as an unchecked CNum   circuit/r1cs/num.rs:26-106,
#     circuit/r1cs/bool.rs:13-106
#   * c_into_bits_le, c_comp_constant, c_into_bits_le_strict                    circuit/bitify.rs:9-112
#   * c_mux3                                                                    circuit/mux.rs:8-32
#   * CEdwardsPoint / CMontgomeryPoint gadgets, fixed- and variable-base mul     circuit/ecc.rs:24-283
# Pinned by the reference's README benchmark table (README.md:48-54): ecmul 254 bits = 2,296 gates,
# ecmul_const 254 bits = 513 gates, poseidon (4, 8, 54) = 255 gates (tests/test_frontend.py).  The README's
# whole-circuit figure (3,860) and its "oncurve+subgroup check = 19" row do not match the gadget code at
# this commit (see test_eddsa_circuit_shape for the count the source gives); no reference test proves this
# circuit, so the gate CONTENT is "restated, unpinned" -- every gate is checked to hold on the witness.
# ====================================================================================================
import hashlib

FS = 2736030358979909402780800718157159386076813972158567259200215660948447373041   # engines/bn256/mod.rs:33
FS_BITS = FS.bit_length()
R_BITS = R.bit_length()


def _finv(a):
    return pow(a % R, -1, R)


def fr_sqrt(a):
    """Tonelli-Shanks in Fr; None for a non-residue.  Which root comes back never reaches a witness:
    every caller fixes the sign afterwards (parity rule or cofactor check)."""
    a %= R
    if a == 0:
        return 0
    if pow(a, (R - 1) // 2, R) != 1:
        return None
    q, s = R - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (R - 1) // 2, R) != R - 1:
        z += 1
    m, c, t, r = s, pow(z, q, R), pow(a, q, R), pow(a, (q + 1) // 2, R)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % R
            i += 1
        b = pow(c, 1 << (m - i - 1), R)
        m, c = i, b * b % R
        t, r = t * c % R, r * b % R
    return r


def ed_add(p, q, d):
    """Unified addition on y^2 - x^2 = 1 + d x^2 y^2 (native/ecc.rs:303-330, in affine form)."""
    (x1, y1), (x2, y2) = p, q
    k = d * x1 * x2 * y1 * y2 % R
    return ((x1 * y2 + y1 * x2) * _finv(1 + k) % R, (y1 * y2 + x1 * x2) * _finv(1 - k) % R)


def ed_mul(p, k, d):
    """native/ecc.rs:336-352"""
    res = (0, 1)
    for i in reversed(range(max(k.bit_length(), 1))):
        res = ed_add(res, res, d)
        if (k >> i) & 1:
            res = ed_add(res, p, d)
    return res


def ed_into_montgomery(p):
    """native/ecc.rs:183-199; None for the neutral element."""
    x, y = p
    if x == 0:
        return None if y == 1 else (0, 0)
    mx = (1 + y) * _finv(1 - y) % R
    return (mx, mx * _finv(x) % R)


def mont_into_edwards(p):
    """native/ecc.rs:216-227"""
    x, y = p
    if x == 0:
        return (0, R - 1)
    return (x * _finv(y) % R, (x - 1) * _finv(x + 1) % R)


class JubJubBN256:
    """engines/bn256/mod.rs:48-72"""

    def __init__(self):
        self.d = (-168696) * _finv(168700) % R
        self.a = 2 * (1 - self.d) * _finv(1 + self.d) % R
        self.b = (-4) * _finv(1 + self.d) % R
        self.u = 337401
        t = SeedboxChaCha20(b"edwards_g").gen_fr()
        self.g = self.from_scalar_raw(t)

    def _g(self, x):
        return (x * x * (x + self.a) + x) * _finv(self.b) % R

    def from_scalar_raw(self, t):
        """native/ecc.rs:104-135"""
        t2g1 = t * t * self.u % R
        x2 = (-_finv(self.a)) * (1 + _finv(t2g1)) % R
        y = fr_sqrt(self._g(x2))
        if y is not None:
            mx, my = x2, y
        else:
            mx = x2 * t2g1 % R
            my = fr_sqrt(self._g(mx))
            assert my is not None
        if (my * t % R) & 1:
            my = (-my) % R
        p = mont_into_edwards((mx, my))
        for _ in range(3):
            p = ed_add(p, p, self.d)
        return p

    def in_curve(self, p):
        x2, y2 = p[0] * p[0] % R, p[1] * p[1] % R
        return (y2 - x2) % R == (1 + self.d * x2 * y2) % R

    def subgroup_decompress(self, x):
        """native/ecc.rs:68-91"""
        x2 = x * x % R
        y = fr_sqrt((x2 + 1) * _finv(1 - self.d * x2) % R)
        if y is None:
            return None
        lx, ly = ed_mul((x, y), FS, self.d)
        if lx != 0:
            return None
        return (x, y) if ly == 1 else (x, (-y) % R)


def _limbs_le(v):
    return b"".join(struct.pack("<Q", (v >> (64 * i)) & MASK64) for i in range(4))


def eddsaposeidon_sign(sk, m, pparams: PoseidonParams, jj: JubJubBN256):
    """native/eddsaposeidon.rs:14-52 -> (s in Fs, r_x in Fr)"""
    h = hashlib.blake2s(_limbs_le(sk) + _limbs_le(m), digest_size=32, person=b"__fawkes").digest()
    rho = int.from_bytes(h, "little") % FS
    r_x = ed_mul(jj.g, rho, jj.d)[0]
    a_x = ed_mul(jj.g, sk, jj.d)[0]
    s = (rho + (poseidon([r_x, a_x, m], pparams) % FS) * sk) % FS
    return s, r_x


def eddsaposeidon_verify(s, r, a, m, pparams, jj) -> bool:
    """native/eddsaposeidon.rs:54-79"""
    p_a, p_r = jj.subgroup_decompress(a), jj.subgroup_decompress(r)
    if p_a is None or p_r is None:
        return False
    ha = ed_mul(p_a, poseidon([r, a, m], pparams) % FS, jj.d)
    return ed_mul(jj.g, s, jj.d) == ed_add(ha, p_r, jj.d)


# ------------------------------------------------------------------ CNum gadgets (free functions) ---
def c_neg(a: CNum) -> CNum:
    return a.scale(R - 1)


def c_const_of(a: CNum):
    return a.as_const()


def c_assert_const(a: CNum, value):
    """num.rs:141-147 / bool.rs:73-79: enforce(self, 1, const)"""
    a.cs.enforce(a, a.cs.const(1), a.cs.const(value))


def c_div_unchecked(a: CNum, b: CNum) -> CNum:
    """num.rs:36-46"""
    bc = b.as_const()
    if bc is not None:
        return a.scale(_finv(bc))
    out = a.cs.alloc(a.value * _finv(b.value) % R)
    a.cs.enforce(out, b, a)
    return out


def c_is_zero(a: CNum) -> CNum:
    """num.rs:64-78"""
    c = a.as_const()
    if c is not None:
        return a.cs.const(1 if c == 0 else 0)
    inv = a.cs.alloc(_finv(a.value) if a.value else 0)
    res = c_neg(inv) * a + 1
    c_assert_const(res * a, 0)
    return res


def c_switch(bit: CNum, if_true: CNum, if_false: CNum) -> CNum:
    """num.rs:149-159: bit ? if_true : if_false (a constant bit picks a side for free)"""
    b = bit.as_const()
    if b is not None:
        return if_true if b == 1 else if_false
    return if_false + (if_true - if_false) * bit


def c_alloc_bool(cs: BuildCS, bit) -> CNum:
    b = cs.alloc(1 if bit else 0)
    b.assert_bit()
    return b


def c_into_bits_le(signal: CNum, limit: int):
    """circuit/bitify.rs:9-47"""
    cs = signal.cs
    c = signal.as_const()
    if c is not None:
        assert c >> limit == 0
        return [cs.const((c >> i) & 1) for i in range(limit)]
    remained = signal
    bits = [cs.const(0)] * limit
    k = 1
    for i in range(1, limit):
        k = k * 2 % R
        s = c_alloc_bool(cs, (signal.value >> i) & 1)
        remained = remained - s.scale(k)
        bits[i] = s
    remained.assert_bit()                                     # to_bool() = CBool::new
    bits[0] = remained
    return bits


def c_comp_constant(signal, ct: int) -> CNum:
    """circuit/bitify.rs:60-105: signal (bits, LE) > ct"""
    siglen = len(signal)
    cs = signal[0].cs
    c_false = cs.const(0)
    if ct >> siglen:
        return c_false
    nsteps = (siglen + 1) >> 1
    assert nsteps + 1 < R_BITS
    sig = list(signal) + [c_false] * (2 * nsteps - siglen)
    k = 1
    acc = cs.const(0)
    for j in range(nsteps):
        ct_l, ct_u = (ct >> (2 * j)) & 1, (ct >> (2 * j + 1)) & 1
        sig_l, sig_u = sig[2 * j], sig[2 * j + 1]
        sig_lu = sig_l * sig_u
        if (ct_l, ct_u) == (0, 0):
            term = sig_l + sig_u - sig_lu
        elif (ct_l, ct_u) == (1, 0):
            term = sig_l + sig_u.scale(2) - sig_lu - 1
        elif (ct_l, ct_u) == (0, 1):
            term = sig_lu + sig_u - 1
        else:
            term = sig_lu - 1
        acc = acc + term.scale(k)
        k = k * 2 % R
    acc = acc + (k - 1)
    return c_into_bits_le(acc, nsteps + 1)[nsteps]


def c_into_bits_le_strict(signal: CNum):
    """circuit/bitify.rs:107-112"""
    bits = c_into_bits_le(signal, R_BITS)
    c_assert_const(c_comp_constant(bits, R - 1), 0)
    return bits


def c_mux3(s, c):
    """circuit/mux.rs:8-32: c[i][s0 + 2 s1 + 4 s2] for every column i"""
    assert len(s) == 3 and all(len(col) == 8 for col in c)
    s10 = s[0] * s[1]
    res = []
    for col in c:
        a210 = s10.scale(col[7] - col[6] - col[5] + col[4] - col[3] + col[2] + col[1] - col[0])
        a21 = s[1].scale(col[6] - col[4] - col[2] + col[0])
        a20 = s[0].scale(col[5] - col[4] - col[1] + col[0])
        a2 = (col[4] - col[0]) % R
        a10 = s10.scale(col[3] - col[2] - col[1] + col[0])
        a1 = s[1].scale(col[2] - col[0])
        a0 = s[0].scale(col[1] - col[0])
        res.append((a210 + a21 + a20 + a2) * s[2] + a10 + a1 + a0 + col[0])
    return res


class CMontgomeryPoint:
    """circuit/ecc.rs:244-283"""

    def __init__(self, x: CNum, y: CNum):
        self.x, self.y = x, y

    def double(self, jj):
        x2 = self.x * self.x
        l = c_div_unchecked(x2.scale(3) + self.x.scale(2 * jj.a) + 1, self.y.scale(2 * jj.b))
        b_l2 = (l * l).scale(jj.b)
        return CMontgomeryPoint(b_l2 - jj.a - self.x.scale(2),
                                l * (self.x.scale(3) + jj.a - b_l2) - self.y)

    def add(self, p, jj):
        l = c_div_unchecked(p.y - self.y, p.x - self.x)
        b_l2 = (l * l).scale(jj.b)
        return CMontgomeryPoint(b_l2 - jj.a - self.x - p.x,
                                l * (self.x.scale(2) + p.x + jj.a - b_l2) - self.y)

    def into_edwards(self):
        y_is_zero = c_is_zero(self.y)
        return CEdwardsPoint(c_div_unchecked(self.x, self.y + y_is_zero),
                             c_div_unchecked(self.x - 1, self.x + 1))

    def switch(self, bit, if_else):
        return CMontgomeryPoint(c_switch(bit, self.x, if_else.x), c_switch(bit, self.y, if_else.y))


class CEdwardsPoint:
    """circuit/ecc.rs:10-242"""

    def __init__(self, x: CNum, y: CNum):
        self.x, self.y = x, y

    @staticmethod
    def alloc(cs, p):
        return CEdwardsPoint(cs.alloc(p[0]), cs.alloc(p[1]))     # field by field (derive(Signal))

    @staticmethod
    def from_const(cs, p):
        return CEdwardsPoint(cs.const(p[0]), cs.const(p[1]))

    def as_const(self):
        x, y = self.x.as_const(), self.y.as_const()
        return None if x is None or y is None else (x, y)

    def switch(self, bit, if_else):
        return CEdwardsPoint(c_switch(bit, self.x, if_else.x), c_switch(bit, self.y, if_else.y))

    def double(self, jj):
        v = self.x * self.y
        v2 = v * v
        u = (self.x + self.y) * (self.x + self.y)
        return CEdwardsPoint(c_div_unchecked(v.scale(2), v2.scale(jj.d) + 1),
                             c_div_unchecked(u - v.scale(2), c_neg(v2.scale(jj.d)) + 1))

    def mul_by_cofactor(self, jj):
        return self.double(jj).double(jj).double(jj)

    def add(self, p, jj):
        v1 = self.x * p.y
        v2 = p.x * self.y
        v12 = v1 * v2
        u = (self.x + self.y) * (p.x + p.y)
        return CEdwardsPoint(c_div_unchecked(v1 + v2, v12.scale(jj.d) + 1),
                             c_div_unchecked(u - v1 - v2, c_neg(v12.scale(jj.d)) + 1))

    def assert_in_curve(self, jj):
        x2 = self.x * self.x
        y2 = self.y * self.y
        (x2.scale(jj.d) * y2).assert_eq(y2 - x2 - 1)

    def assert_in_subgroup(self, jj):
        """circuit/ecc.rs:56-67"""
        pre = ed_mul((self.x.value, self.y.value), pow(8, -1, FS), jj.d)
        preimage = CEdwardsPoint.alloc(self.x.cs, pre)
        preimage.assert_in_curve(jj)
        p8 = preimage.mul_by_cofactor(jj)
        c_assert_const(self.x - p8.x, 0)
        c_assert_const(self.y - p8.y, 0)

    @staticmethod
    def subgroup_decompress(x: CNum, jj):
        """circuit/ecc.rs:69-80"""
        p = jj.subgroup_decompress(x.value) or jj.g
        preimage = CEdwardsPoint.alloc(x.cs, ed_mul(p, pow(8, -1, FS), jj.d))
        preimage.assert_in_curve(jj)
        p8 = preimage.mul_by_cofactor(jj)
        c_assert_const(x - p8.x, 0)
        return p8

    def into_montgomery(self):
        x = c_div_unchecked(self.y + 1, c_neg(self.y) + 1)
        return CMontgomeryPoint(x, c_div_unchecked(x, self.x))

    def mul(self, bits, jj):
        """circuit/ecc.rs:91-189: 3-bit windows over a table for a constant base, conditional Montgomery
        adds over precomputed doublings otherwise."""
        cs = bits[0].cs
        c_base = self.as_const()
        if c_base is not None:
            if c_base == (0, 1):
                return CEdwardsPoint.from_const(cs, (0, 1))
            all_bits = list(bits) + [cs.const(0)] * ((2 * len(bits)) % 3)
            nwindows = len(all_bits) // 3
            acc, base = (0, R - 1), c_base
            for _ in range(nwindows):
                acc = ed_add(acc, base, jj.d)
                for _ in range(3):
                    base = ed_add(base, base, jj.d)
            mp = ed_into_montgomery(((-acc[0]) % R, acc[1]))
            cacc = CMontgomeryPoint(cs.const(mp[0]), cs.const(mp[1]))
            base = c_base
            for i in range(nwindows):
                q, xs, ys = base, [], []
                for _ in range(8):
                    mx, my = ed_into_montgomery(q)
                    xs.append(mx)
                    ys.append(my)
                    q = ed_add(q, base, jj.d)
                rx, ry = c_mux3(all_bits[3 * i:3 * i + 3], [xs, ys])
                cacc = cacc.add(CMontgomeryPoint(rx, ry), jj)
                for _ in range(3):
                    base = ed_add(base, base, jj.d)
            res = cacc.into_edwards()
            return CEdwardsPoint(c_neg(res.x), c_neg(res.y))
        base_is_zero = c_is_zero(self.x)
        dummy = CEdwardsPoint.from_const(cs, jj.g)
        bp = dummy.switch(base_is_zero, self).into_montgomery()
        exponents = [bp]
        for _ in range(1, len(bits)):
            bp = bp.double(jj)
            exponents.append(bp)
        empty = CMontgomeryPoint(cs.const(0), cs.const(0))
        acc = empty
        for i in range(len(bits)):
            acc = acc.add(exponents[i], jj).switch(bits[i], acc)
        acc = empty.switch(base_is_zero, acc)
        res = acc.into_edwards()
        return CEdwardsPoint(c_neg(res.x), c_neg(res.y))


def c_eddsaposeidon_verify(s: CNum, r: CNum, a: CNum, m: CNum, pparams: PoseidonParams, jj: JubJubBN256) -> CNum:
    """circuit/eddsaposeidon.rs:16-47 -> CBool (as an unchecked CNum)"""
    assert R_BITS > FS_BITS
    cs = s.cs
    p_a = CEdwardsPoint.subgroup_decompress(a, jj)
    p_r = CEdwardsPoint.subgroup_decompress(r, jj)
    h = c_poseidon([r, a, m], pparams)
    h_bits = c_into_bits_le_strict(h)
    ha = p_a.mul(h_bits, jj)
    s_bits = c_into_bits_le(s, FS_BITS)
    c_assert_const(c_comp_constant(s_bits, FS - 1), 0)
    sb = CEdwardsPoint.from_const(cs, jj.g).mul(s_bits, jj)
    ha_plus_r = ha.add(p_r, jj)
    return c_is_zero(ha_plus_r.x - sb.x)


def eddsa_circuit(sk: int, m: int, pparams: PoseidonParams = None, jj: JubJubBN256 = None, forge: bool = False):
    """configs[1] as the prover would see it (the reference has no Groth16 test of this circuit, so the
    closure is ours: Pub = the message m, Sec = (s, r, a); `c_eddsaposeidon_verify(..).assert_const(&true)`).
    Returns (gates, values_input, values_aux)."""
    pparams = pparams or PoseidonParams(4, 8, 54)
    jj = jj or JubJubBN256()
    s, r_x = eddsaposeidon_sign(sk, m, pparams, jj)
    a_x = ed_mul(jj.g, sk, jj.d)[0]
    assert eddsaposeidon_verify(s, r_x, a_x, m, pparams, jj)
    cs = BuildCS()
    c_m = cs.alloc(m)
    cs.inputize(c_m)
    c_s, c_r, c_a = cs.alloc(s), cs.alloc(r_x), cs.alloc(a_x)
    ok = c_eddsaposeidon_verify(c_s, c_r, c_a, c_m, pparams, jj)
    assert ok.value == 1
    c_assert_const(ok, 1)
    return cs.gates, cs.inputs, cs.aux
