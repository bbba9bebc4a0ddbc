"""BN254 field / curve / pairing arithmetic on Python integers.

TEST INFRASTRUCTURE ONLY.  This file is part of the CPU oracle: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.  The
product path (fawkes-crypto_b200/) never does.

PARITY STATUS: the reference (zeropoolnetwork/fawkes-crypto) delegates all of
this arithmetic to crates that are NOT in /root/reference
(fawkes-crypto-bellman_ce 0.3.5, fawkes-crypto-pairing_ce 0.18.1, ff_ce 0.7.1;
Cargo.lock:413-436,495-504) and holds no golden proof vectors -> proof-level
"parity unpinned".  What IS pinned against the reference's own tests:
the generic Montgomery field code (class MontField below) replays the decimal
known-answer tests of ff-uint/tests/ff-uint_tests.rs:35-153.

What each part follows:
  * MontField          ff-uint_derive/src/lib.rs:221-405 (constants),
                       :434-490 (mont_reduce), :578-623 (mul), :836-862
                       (add/sub/neg), :864-917 (inverse)
  * moduli / generator fawkes-crypto/src/engines/bn256/mod.rs:8-26
  * curve, Fq2 tower, optimal-ate pairing: published BN254 (alt_bn128)
    algorithm as used by pairing_ce bn256 (y^2=x^3+3, Fq2=Fq[u]/(u^2+1),
    twist y^2=x^3+3/(9+u), BN x=4965661367192848881).
"""
from __future__ import annotations

# --------------------------------------------------------------------------
# moduli (fawkes-crypto/src/engines/bn256/mod.rs:13,23)
# --------------------------------------------------------------------------
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
FR_GENERATOR = 7            # engines/bn256/mod.rs:24
BN_X = 4965661367192848881
MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------
# Generic limb-level Montgomery field, restating ff-uint_derive's generated code
# --------------------------------------------------------------------------
class MontField:
    """Limb-exact restatement of `construct_primefield_params!` output."""

    def __init__(self, modulus: int, generator: int, limbs: int | None = None):
        self.m = modulus
        # ff-uint_derive/src/lib.rs:79-89: limbs so that 2*modulus fits
        if limbs is None:
            limbs = 1
            while (2 * modulus) >> (64 * limbs):
                limbs += 1
        self.limbs = limbs
        self.bits = modulus.bit_length()
        self.shave = 64 * limbs - self.bits                      # lib.rs:234
        self.Rint = (1 << (64 * limbs)) % modulus                # lib.rs:237
        s, t = 0, modulus - 1
        while t % 2 == 0:                                        # lib.rs:240-246
            t >>= 1
            s += 1
        self.S, self.T = s, t
        self.root_of_unity = pow(generator, t, modulus)          # canonical value
        self.generator = generator
        self.R2int = (self.Rint * self.Rint) % modulus
        inv = 1                                                  # lib.rs:360-366
        for _ in range(63):
            inv = (inv * inv) & MASK64
            inv = (inv * (modulus & MASK64)) & MASK64
        self.INV = (-inv) & MASK64
        self.mod_limbs = self.to_limbs(modulus)

    # -- limb helpers ------------------------------------------------------
    def to_limbs(self, x: int):
        return [(x >> (64 * i)) & MASK64 for i in range(self.limbs)]

    @staticmethod
    def from_limbs(l):
        return sum(v << (64 * i) for i, v in enumerate(l))

    # -- Montgomery core (operates on Montgomery-form integers < m) --------
    def mont_reduce(self, t):
        """t: list of 2*limbs u64 limbs -> reduced integer (lib.rs:434-490)."""
        n = self.limbs
        t = list(t)
        carry2 = 0
        for i in range(n):
            k = (t[i] * self.INV) & MASK64
            carry = 0
            v = t[i] + k * self.mod_limbs[0] + carry
            carry = v >> 64
            for j in range(1, n):
                v = t[i + j] + k * self.mod_limbs[j] + carry
                t[i + j] = v & MASK64
                carry = v >> 64
            v = t[i + n] + carry2 + carry
            t[i + n] = v & MASK64
            carry2 = v >> 64
        res = self.from_limbs(t[n:2 * n])
        if res >= self.m:                                        # reduced(), lib.rs:947-954
            res -= self.m
        return res

    def mont_mul(self, a: int, b: int) -> int:
        """Schoolbook product then mont_reduce (lib.rs:578-623)."""
        n = self.limbs
        al, bl = self.to_limbs(a), self.to_limbs(b)
        t = [0] * (2 * n)
        for i in range(n):
            carry = 0
            for j in range(n):
                v = t[i + j] + al[i] * bl[j] + carry
                t[i + j] = v & MASK64
                carry = v >> 64
            t[i + n] = carry
        return self.mont_reduce(t)

    def to_mont(self, x: int) -> int:      # from_uint: x*R2 (lib.rs:776-783)
        assert 0 <= x < self.m
        return self.mont_mul(x, self.R2int)

    def from_mont(self, x: int) -> int:    # to_uint: mont_reduce(x,0..) (lib.rs:785-787)
        return self.mont_reduce(self.to_limbs(x) + [0] * self.limbs)

    def add(self, a, b):                   # lib.rs:836-839
        r = a + b
        return r - self.m if r >= self.m else r

    def sub(self, a, b):                   # lib.rs:846-853
        return a - b if a >= b else a + (self.m - b)

    def neg(self, a):                      # lib.rs:855-862
        return 0 if a == 0 else self.m - a

    def inverse(self, a):
        """Binary extended Euclid of lib.rs:864-917 on Montgomery values."""
        if a == 0:
            return None
        m = self.m
        u, v = a, m
        b, c = self.R2int, 0
        while u != 1 and v != 1:
            while u % 2 == 0:
                u >>= 1
                b = b >> 1 if b % 2 == 0 else (b + m) >> 1
            while v % 2 == 0:
                v >>= 1
                c = c >> 1 if c % 2 == 0 else (c + m) >> 1
            if v < u:
                u -= v
                b = self.sub(b, c)
            else:
                v -= u
                c = self.sub(c, b)
        return b if u == 1 else c

    def pow(self, a, e):
        res = self.Rint
        for i in reversed(range(e.bit_length())):
            res = self.mont_mul(res, res)
            if (e >> i) & 1:
                res = self.mont_mul(res, a)
        return res

    def legendre(self, a):
        s = self.pow(a, (self.m - 1) >> 1)
        if s == 0:
            return 0
        return 1 if s == self.Rint else -1

    def sqrt(self, a):
        """Canonical-in / canonical-out square root via Tonelli-Shanks or the
        p=3 mod 4 shortcut (lib.rs:271-352); operates on Montgomery values."""
        m = self.m
        if self.legendre(a) == 0:
            return a
        if self.legendre(a) != 1:
            return None
        if m % 4 == 3:
            r = self.pow(a, (m + 1) >> 2)
            return r
        # Tonelli-Shanks, same structure as lib.rs:300-352
        c = self.to_mont(self.root_of_unity)
        r = self.pow(a, (self.T + 1) >> 1)
        t = self.pow(a, self.T)
        mm = self.S
        while t != self.Rint:
            i = 1
            t2i = self.mont_mul(t, t)
            while t2i != self.Rint:
                t2i = self.mont_mul(t2i, t2i)
                i += 1
            for _ in range(mm - i - 1):
                c = self.mont_mul(c, c)
            r = self.mont_mul(r, c)
            c = self.mont_mul(c, c)
            t = self.mont_mul(t, c)
            mm = i
        return r


FR = MontField(R, FR_GENERATOR)
FQ = MontField(P, 3)   # GENERATOR for Fq irrelevant to the prover path

FR_S = FR.S                            # 28
FR_ROOT_OF_UNITY = FR.root_of_unity    # canonical 7^t
R_MONT = FR.Rint                       # 2^256 mod r
P_MONT = FQ.Rint                       # 2^256 mod p


def fr_to_mont(x): return (x * R_MONT) % R
def fr_from_mont(x): return (x * pow(R_MONT, -1, R)) % R
def fq_to_mont(x): return (x * P_MONT) % P
def fq_from_mont(x): return (x * pow(P_MONT, -1, P)) % P


def limbs4(x: int):
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


# --------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1) as tuples (c0, c1)
# --------------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a):
    return ((a[0] + a[1]) * (a[0] - a[1]) % P, 2 * a[0] * a[1] % P)
def f2_muls(a, k): return (a[0] * k % P, a[1] * k % P)
def f2_conj(a): return (a[0], (-a[1]) % P)
def f2_inv(a):
    t = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * t % P, (-a[1]) * t % P)
def f2_pow(a, e):
    res = (1, 0)
    for i in reversed(range(e.bit_length())):
        res = f2_sqr(res)
        if (e >> i) & 1:
            res = f2_mul(res, a)
    return res

F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (9, 1)                                       # Fq6 non-residue 9+u
B1 = 3                                            # G1: y^2 = x^3 + 3
B2 = f2_mul((3, 0), f2_inv(XI))                   # G2: y^2 = x^3 + 3/(9+u)

# --------------------------------------------------------------------------
# Curve groups: a generic short-Weierstrass (a=0) implementation over an
# abstract field given by an ops table.  Points: None = infinity, else (x, y)
# affine; Jacobian (X, Y, Z) with Z == zero meaning infinity.
# --------------------------------------------------------------------------
class _Ops:
    def __init__(self, add, sub, mul, sqr, neg, inv, zero, one, b):
        self.add, self.sub, self.mul, self.sqr = add, sub, mul, sqr
        self.neg, self.inv, self.zero, self.one, self.b = neg, inv, zero, one, b

OPS1 = _Ops(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P,
            lambda a, b: a * b % P, lambda a: a * a % P, lambda a: (-a) % P,
            lambda a: pow(a, -1, P), 0, 1, B1)
OPS2 = _Ops(f2_add, f2_sub, f2_mul, f2_sqr, f2_neg, f2_inv, F2_ZERO, F2_ONE, B2)


def on_curve(o: _Ops, pt):
    if pt is None:
        return True
    x, y = pt
    return o.sqr(y) == o.add(o.mul(o.sqr(x), x), o.b)


def to_jac(o, pt):
    return (o.one, o.one, o.zero) if pt is None else (pt[0], pt[1], o.one)


def to_affine(o, j):
    X, Y, Z = j
    if Z == o.zero:
        return None
    zi = o.inv(Z)
    zi2 = o.sqr(zi)
    return (o.mul(X, zi2), o.mul(Y, o.mul(zi2, zi)))


def jac_double(o, j):
    X, Y, Z = j
    if Z == o.zero:
        return j
    A = o.sqr(X); B = o.sqr(Y); C = o.sqr(B)
    D = o.sub(o.sub(o.sqr(o.add(X, B)), A), C); D = o.add(D, D)
    E = o.add(o.add(A, A), A)
    F = o.sqr(E)
    X3 = o.sub(F, o.add(D, D))
    C8 = o.add(C, C); C8 = o.add(C8, C8); C8 = o.add(C8, C8)
    Y3 = o.sub(o.mul(E, o.sub(D, X3)), C8)
    Z3 = o.mul(o.add(Y, Y), Z)
    return (X3, Y3, Z3)


def jac_add(o, p, q):
    if p[2] == o.zero:
        return q
    if q[2] == o.zero:
        return p
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    Z1Z1 = o.sqr(Z1); Z2Z2 = o.sqr(Z2)
    U1 = o.mul(X1, Z2Z2); U2 = o.mul(X2, Z1Z1)
    S1 = o.mul(o.mul(Y1, Z2), Z2Z2); S2 = o.mul(o.mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return jac_double(o, p)
        return (o.one, o.one, o.zero)
    H = o.sub(U2, U1)
    Rr = o.sub(S2, S1)
    HH = o.sqr(H); HHH = o.mul(H, HH)
    V = o.mul(U1, HH)
    X3 = o.sub(o.sub(o.sqr(Rr), HHH), o.add(V, V))
    Y3 = o.sub(o.mul(Rr, o.sub(V, X3)), o.mul(S1, HHH))
    Z3 = o.mul(o.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def jac_neg(o, p):
    return (p[0], o.neg(p[1]), p[2])


def jac_mul(o, p, k: int):
    if k < 0:
        return jac_mul(o, jac_neg(o, p), -k)
    acc = (o.one, o.one, o.zero)
    for i in reversed(range(k.bit_length())):
        acc = jac_double(o, acc)
        if (k >> i) & 1:
            acc = jac_add(o, acc, p)
    return acc


def pt_add(o, a, b):
    return to_affine(o, jac_add(o, to_jac(o, a), to_jac(o, b)))


def pt_mul(o, a, k):
    return to_affine(o, jac_mul(o, to_jac(o, a), k))


def pt_neg(o, a):
    return None if a is None else (a[0], o.neg(a[1]))


def batch_to_affine(o, js):
    """Montgomery-trick normalisation of many Jacobian points."""
    prods, acc = [], o.one
    for X, Y, Z in js:
        if Z != o.zero:
            acc = o.mul(acc, Z)
        prods.append(acc)
    inv = o.inv(acc)
    out = [None] * len(js)
    for i in reversed(range(len(js))):
        X, Y, Z = js[i]
        if Z == o.zero:
            continue
        prev = prods[i - 1] if i > 0 else o.one
        zi = o.mul(inv, prev)
        inv = o.mul(inv, Z)
        zi2 = o.sqr(zi)
        out[i] = (o.mul(X, zi2), o.mul(Y, o.mul(zi2, zi)))
    return out


class FixedBase:
    """Windowed fixed-base scalar multiplication table (host-side setup helper)."""

    def __init__(self, o, base, window=8, bits=256):
        self.o, self.w = o, window
        self.nwin = (bits + window - 1) // window
        self.table = []
        cur = to_jac(o, base)
        for _ in range(self.nwin):
            row = [(o.one, o.one, o.zero)]
            for _ in range((1 << window) - 1):
                row.append(jac_add(o, row[-1], cur))
            row = [to_jac(o, p) for p in batch_to_affine(o, row)]
            self.table.append(row)
            for _ in range(window):
                cur = jac_double(o, cur)

    def mul_jac(self, k: int):
        o = self.o
        acc = (o.one, o.one, o.zero)
        mask = (1 << self.w) - 1
        i = 0
        while k:
            d = k & mask
            if d:
                acc = jac_add(o, acc, self.table[i][d])
            k >>= self.w
            i += 1
        return acc

    def mul_many(self, ks):
        return batch_to_affine(self.o, [self.mul_jac(k) for k in ks])


G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)

# --------------------------------------------------------------------------
# Fq12 = Fq2[w]/(w^6 - xi): list of 6 Fq2 coefficients.  Used by the pairing.
# --------------------------------------------------------------------------
def f12_one():
    return [F2_ONE] + [F2_ZERO] * 5


def f12_mul(a, b):
    t = [[0, 0] for _ in range(11)]
    for i in range(6):
        ai = a[i]
        if ai == F2_ZERO:
            continue
        a0, a1 = ai
        for j in range(6):
            b0, b1 = b[j]
            if b0 == 0 and b1 == 0:
                continue
            t[i + j][0] += a0 * b0 - a1 * b1
            t[i + j][1] += a0 * b1 + a1 * b0
    res = []
    for k in range(6):
        c0, c1 = t[k]
        if k + 6 < 11:
            h0, h1 = t[k + 6]
            # (h0 + h1 u)(9 + u) = 9h0 - h1 + (h0 + 9h1) u
            c0 += 9 * h0 - h1
            c1 += h0 + 9 * h1
        res.append((c0 % P, c1 % P))
    return res


def f12_sqr(a):
    return f12_mul(a, a)


def f12_pow(a, e):
    res = f12_one()
    for i in reversed(range(e.bit_length())):
        res = f12_sqr(res)
        if (e >> i) & 1:
            res = f12_mul(res, a)
    return res


def f12_conj(a):
    """a^(p^6): w -> -w."""
    return [a[i] if i % 2 == 0 else f2_neg(a[i]) for i in range(6)]


_FROB_W = [f2_pow(XI, i * (P - 1) // 6) for i in range(6)]   # w^(i(p-1))


def f12_frob(a):
    """a^p: conjugate each Fq2 coefficient, multiply by xi^(i(p-1)/6)."""
    return [f2_mul(f2_conj(a[i]), _FROB_W[i]) for i in range(6)]


def f12_inv(a):
    # a^-1 = a^(p^12 - 2); cheaper: use norm to Fq6-ish via conj trick:
    # a * conj(a) lies in the subfield fixed by w->-w (even powers only), i.e.
    # Fq6 = Fq2[v]/(v^3 - xi) with v = w^2.  Invert there by cubic formulas.
    n = f12_mul(a, f12_conj(a))
    c0, c1, c2 = n[0], n[2], n[4]
    # inverse in Fq2[v]/(v^3 - xi)
    t0 = f2_sub(f2_sqr(c0), f2_mul(XI, f2_mul(c1, c2)))
    t1 = f2_sub(f2_mul(XI, f2_sqr(c2)), f2_mul(c0, c1))
    t2 = f2_sub(f2_sqr(c1), f2_mul(c0, c2))
    d = f2_add(f2_mul(c0, t0), f2_mul(XI, f2_add(f2_mul(c2, t1), f2_mul(c1, t2))))
    di = f2_inv(d)
    ni = [f2_mul(t0, di), F2_ZERO, f2_mul(t1, di), F2_ZERO, f2_mul(t2, di), F2_ZERO]
    return f12_mul(f12_conj(a), ni)


ATE_LOOP = 6 * BN_X + 2
_GAMMA_X = f2_pow(XI, (P - 1) // 3)
_GAMMA_Y = f2_pow(XI, (P - 1) // 2)


def _twist_frob(q):
    return (f2_mul(f2_conj(q[0]), _GAMMA_X), f2_mul(f2_conj(q[1]), _GAMMA_Y))


def _line(T, Q, Pt):
    """Line through twist points T,Q (tangent if equal) evaluated at G1 point
    Pt, as a sparse Fq12 element; also returns T+Q (affine on the twist)."""
    xP, yP = Pt
    if T[0] == Q[0] and T[1] == Q[1]:
        lam = f2_mul(f2_muls(f2_sqr(T[0]), 3), f2_inv(f2_muls(T[1], 2)))
    else:
        lam = f2_mul(f2_sub(Q[1], T[1]), f2_inv(f2_sub(Q[0], T[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), T[0]), Q[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(T[0], x3)), T[1])
    l = [F2_ZERO] * 6
    l[0] = (yP % P, 0)
    l[1] = f2_neg(f2_muls(lam, xP))
    l[3] = f2_sub(f2_mul(lam, T[0]), T[1])
    return l, (x3, y3)


def miller_loop(Pt, Q):
    """Optimal-ate Miller loop f_{6x+2,Q}(P) with the two Frobenius lines."""
    if Pt is None or Q is None:
        return f12_one()
    f = f12_one()
    T = Q
    for i in reversed(range(ATE_LOOP.bit_length() - 1)):
        l, T = _line(T, T, Pt)
        f = f12_mul(f12_sqr(f), l)
        if (ATE_LOOP >> i) & 1:
            l, T = _line(T, Q, Pt)
            f = f12_mul(f, l)
    Q1 = _twist_frob(Q)
    Q2 = _twist_frob(Q1)
    nQ2 = (Q2[0], f2_neg(Q2[1]))
    l, T = _line(T, Q1, Pt)
    f = f12_mul(f, l)
    l, T = _line(T, nQ2, Pt)
    f = f12_mul(f, l)
    return f


def final_exp(f):
    # easy part: f^((p^6-1)(p^2+1)), then hard part by plain exponentiation
    f1 = f12_mul(f12_conj(f), f12_inv(f))
    f2 = f12_mul(f12_frob(f12_frob(f1)), f1)
    hard = (P ** 4 - P ** 2 + 1) // R
    return f12_pow(f2, hard)


def pairing(Pt, Q):
    return final_exp(miller_loop(Pt, Q))


def pairing_product_is_one(pairs):
    f = f12_one()
    for Pt, Q in pairs:
        f = f12_mul(f, miller_loop(Pt, Q))
    return final_exp(f) == f12_one()
