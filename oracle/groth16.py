"""Groth16 setup / prove / verify on Python integers -- the truth oracle.

TEST INFRASTRUCTURE ONLY (see oracle/bn254.py header for the import rule).

PARITY STATUS: "parity unpinned" at proof level.  The reference's prover body is
`bellman::groth16::create_random_proof` (fawkes-crypto/src/backend/
bellman_groth16/prover.rs:80) which lives in the un-vendored crate
fawkes-crypto-bellman_ce 0.3.5 (Cargo.lock:413-425); the reference holds no
golden proofs (fawkes-crypto/tests/bellman_groth16.rs:43-46 only asserts
verify==true).  This file restates bellman_ce's published algorithm
 and is anchored by (i) prove -> independent pairing verify,
(ii) the trapdoor scalar identity A == (alpha + A(tau) + r*delta)*G1 etc.

Row / variable conventions follow the in-repo call sites:
  * gate = (A, B, C) lists of (coeff, Index)     circuit/r1cs/cs.rs:22-26
  * variables_input[0] = ONE, inputs then aux    backend/bellman_groth16/mod.rs:61-102
  * one extra row `input_i * 0 = 0` per input    bellman generator/prover [App. C.1]
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import bn254 as bn
from .bn254 import R, OPS1, OPS2, G1_GEN, G2_GEN

INPUT, AUX = 0, 1          # Index tag, circuit/r1cs/lc.rs:144-149
Term = Tuple[int, Tuple[int, int]]     # (canonical coeff, (tag, idx))
Gate = Tuple[List[Term], List[Term], List[Term]]


# --------------------------------------------------------------------------
# Evaluation domain (bellman_ce domain.rs restated)
# --------------------------------------------------------------------------
def domain_params(n_rows: int):
    m, exp = 1, 0
    while m < n_rows:
        m *= 2
        exp += 1
        if exp >= bn.FR_S:
            raise ValueError("PolynomialDegreeTooLarge")
    omega = bn.FR_ROOT_OF_UNITY
    for _ in range(exp, bn.FR_S):
        omega = omega * omega % R
    return m, exp, omega


def _bitrev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def fft(a: List[int], omega: int, exp: int) -> None:
    """In-place radix-2 DIT, natural order in and out (serial_fft)."""
    n = len(a)
    assert n == 1 << exp
    for k in range(n):
        rk = _bitrev(k, exp)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(exp):
        w_m = pow(omega, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % R
                a[k + j + m] = (a[k + j] - t) % R
                a[k + j] = (a[k + j] + t) % R
                w = w * w_m % R
        m *= 2


def ifft(a, omega, exp):
    fft(a, pow(omega, -1, R), exp)
    minv = pow(len(a), -1, R)
    for i in range(len(a)):
        a[i] = a[i] * minv % R


def distribute_powers(a, g):
    u = 1
    for i in range(len(a)):
        a[i] = a[i] * u % R
        u = u * g % R


def coset_fft(a, omega, exp):
    distribute_powers(a, bn.FR_GENERATOR)
    fft(a, omega, exp)


def icoset_fft(a, omega, exp):
    ifft(a, omega, exp)
    distribute_powers(a, pow(bn.FR_GENERATOR, -1, R))


def h_coefficients(a: List[int], b: List[int], c: List[int]) -> List[int]:
    """(A*B - C)/Z as m-1 coefficients; a,b,c = row evaluations (len n_rows)."""
    m, exp, omega = domain_params(len(a))
    a = a + [0] * (m - len(a))
    b = b + [0] * (m - len(b))
    c = c + [0] * (m - len(c))
    for v in (a, b, c):
        ifft(v, omega, exp)
        coset_fft(v, omega, exp)
    zinv = pow((pow(bn.FR_GENERATOR, m, R) - 1) % R, -1, R)
    for i in range(m):
        a[i] = (a[i] * b[i] - c[i]) % R * zinv % R
    icoset_fft(a, omega, exp)
    return a[:m - 1]


# --------------------------------------------------------------------------
# R1CS evaluation (bellman ProvingAssignment::enforce/eval restated, App. C.1)
# --------------------------------------------------------------------------
def eval_lc(lc, inputs, aux, in_density=None, aux_density=None):
    acc = 0
    for coeff, (tag, idx) in lc:
        if tag == INPUT:
            v = inputs[idx]
            if in_density is not None:
                in_density[idx] = True
        else:
            v = aux[idx]
            if aux_density is not None:
                aux_density[idx] = True
        acc = (acc + coeff * v) % R
    return acc


@dataclass
class Assignment:
    a: List[int]
    b: List[int]
    c: List[int]
    a_aux_density: List[bool]
    b_in_density: List[bool]
    b_aux_density: List[bool]


def evaluate_r1cs(gates: List[Gate], inputs: List[int], aux: List[int]) -> Assignment:
    n_in, n_aux = len(inputs), len(aux)
    a_aux_d = [False] * n_aux
    b_in_d = [False] * n_in
    b_aux_d = [False] * n_aux
    a, b, c = [], [], []
    for A, B, C in gates:
        a.append(eval_lc(A, inputs, aux, None, a_aux_d))
        b.append(eval_lc(B, inputs, aux, b_in_d, b_aux_d))
        c.append(eval_lc(C, inputs, aux))
    for i in range(n_in):          # input_i * 0 = 0
        a.append(inputs[i])
        b.append(0)
        c.append(0)
    return Assignment(a, b, c, a_aux_d, b_in_d, b_aux_d)


def structural_density(gates: List[Gate], n_in: int, n_aux: int):
    a_aux_d = [False] * n_aux
    b_in_d = [False] * n_in
    b_aux_d = [False] * n_aux
    for A, B, _ in gates:
        for _, (tag, idx) in A:
            if tag == AUX:
                a_aux_d[idx] = True
        for _, (tag, idx) in B:
            if tag == INPUT:
                b_in_d[idx] = True
            else:
                b_aux_d[idx] = True
    return a_aux_d, b_in_d, b_aux_d


# --------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------
@dataclass
class VerifyingKey:
    alpha_g1: tuple
    beta_g1: tuple
    beta_g2: tuple
    gamma_g2: tuple
    delta_g1: tuple
    delta_g2: tuple
    ic: list


@dataclass
class Params:
    vk: VerifyingKey
    h: list
    l: list
    a: list
    b_g1: list
    b_g2: list
    # oracle-only extras: discrete logs of every base w.r.t. the generators
    dlog: dict = field(default_factory=dict)


@dataclass
class Trapdoor:
    alpha: int
    beta: int
    gamma: int
    delta: int
    tau: int


def setup(gates: List[Gate], n_in: int, n_aux: int, td: Trapdoor,
          g1=G1_GEN, g2=G2_GEN, want_points=True) -> Params:
    """bellman generate_parameters restated (App. C.4) with explicit trapdoor."""
    n_rows = len(gates) + n_in
    m, exp, omega = domain_params(n_rows)
    tau = td.tau
    powers = [1] * m
    for i in range(1, m):
        powers[i] = powers[i - 1] * tau % R
    z_tau = (pow(tau, m, R) - 1) % R
    coeff = z_tau * pow(td.delta, -1, R) % R
    h_s = [powers[i] * coeff % R for i in range(m - 1)]
    lag = list(powers)
    ifft(lag, omega, exp)           # Lagrange basis polys at tau

    a_in = [0] * n_in; b_in = [0] * n_in; c_in = [0] * n_in
    a_ax = [0] * n_aux; b_ax = [0] * n_aux; c_ax = [0] * n_aux

    def acc(lc, tin, tax, lj):
        for cf, (tag, idx) in lc:
            if tag == INPUT:
                tin[idx] = (tin[idx] + cf * lj) % R
            else:
                tax[idx] = (tax[idx] + cf * lj) % R

    for j, (A, B, C) in enumerate(gates):
        acc(A, a_in, a_ax, lag[j])
        acc(B, b_in, b_ax, lag[j])
        acc(C, c_in, c_ax, lag[j])
    for i in range(n_in):
        a_in[i] = (a_in[i] + lag[len(gates) + i]) % R

    ginv, dinv = pow(td.gamma, -1, R), pow(td.delta, -1, R)
    ic_s = [(td.beta * a_in[i] + td.alpha * b_in[i] + c_in[i]) % R * ginv % R for i in range(n_in)]
    l_s = [(td.beta * a_ax[i] + td.alpha * b_ax[i] + c_ax[i]) % R * dinv % R for i in range(n_aux)]
    a_s = [x for x in a_in + a_ax if x != 0]
    b_s = [x for x in b_in + b_ax if x != 0]

    dlog = dict(h=h_s, l=l_s, a=a_s, b=b_s, ic=ic_s, alpha=td.alpha, beta=td.beta,
                gamma=td.gamma, delta=td.delta, a_full=a_in + a_ax, b_full=b_in + b_ax)
    if not want_points:
        return Params(None, [], [], [], [], [], dlog)
    fb1 = bn.FixedBase(OPS1, g1)
    fb2 = bn.FixedBase(OPS2, g2)
    al1, be1, de1 = fb1.mul_many([td.alpha, td.beta, td.delta])
    be2, ga2, de2 = fb2.mul_many([td.beta, td.gamma, td.delta])
    vk = VerifyingKey(al1, be1, be2, ga2, de1, de2, fb1.mul_many(ic_s))
    return Params(vk, fb1.mul_many(h_s), fb1.mul_many(l_s), fb1.mul_many(a_s),
                  fb1.mul_many(b_s), fb2.mul_many(b_s), dlog)


# --------------------------------------------------------------------------
# Multi-exponentiation: windowed bucket method, unsigned digits (independent of
# the GPU's signed-digit layout on purpose)
# --------------------------------------------------------------------------
def msm(o, bases, scalars, c: Optional[int] = None):
    assert len(bases) == len(scalars)
    n = len(bases)
    if n == 0:
        return (o.one, o.one, o.zero)
    if c is None:
        c = 3 if n < 32 else max(3, min(12, n.bit_length() - 2))
    jb = [bn.to_jac(o, b) for b in bases]
    acc = (o.one, o.one, o.zero)
    nwin = (254 + c - 1) // c
    for w in reversed(range(nwin)):
        for _ in range(c):
            acc = bn.jac_double(o, acc)
        buckets = [None] * ((1 << c) - 1)
        for i in range(n):
            d = (scalars[i] >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = jb[i] if buckets[d - 1] is None else bn.jac_add(o, buckets[d - 1], jb[i])
        run = (o.one, o.one, o.zero)
        tot = (o.one, o.one, o.zero)
        for bkt in reversed(buckets):
            if bkt is not None:
                run = bn.jac_add(o, run, bkt)
            tot = bn.jac_add(o, tot, run)
        acc = bn.jac_add(o, acc, tot)
    return acc


# --------------------------------------------------------------------------
# Prover (bellman create_proof restated, App. C.1-C.5)
# --------------------------------------------------------------------------
@dataclass
class Proof:
    a: tuple
    b: tuple
    c: tuple


def select(values, density):
    return [v for v, d in zip(values, density) if d]


def prove(params: Params, gates: List[Gate], inputs: List[int], aux: List[int],
          r: int, s: int, return_h=False):
    asg = evaluate_r1cs(gates, inputs, aux)
    h = h_coefficients(asg.a, asg.b, asg.c)
    n_in = len(inputs)
    vk = params.vk
    J1 = lambda p: bn.to_jac(OPS1, p)
    J2 = lambda p: bn.to_jac(OPS2, p)
    add1 = lambda x, y: bn.jac_add(OPS1, x, y)
    add2 = lambda x, y: bn.jac_add(OPS2, x, y)

    assert len(params.h) >= len(h)
    h_acc = msm(OPS1, params.h[:len(h)], h)
    l_acc = msm(OPS1, params.l, aux)

    a_aux_sc = select(aux, asg.a_aux_density)
    assert len(params.a) == n_in + len(a_aux_sc), "a query/density mismatch"
    a_ans = add1(msm(OPS1, params.a[:n_in], inputs), msm(OPS1, params.a[n_in:], a_aux_sc))

    b_in_sc = select(inputs, asg.b_in_density)
    b_aux_sc = select(aux, asg.b_aux_density)
    nb_in = len(b_in_sc)
    assert len(params.b_g1) == nb_in + len(b_aux_sc), "b query/density mismatch"
    b1_ans = add1(msm(OPS1, params.b_g1[:nb_in], b_in_sc), msm(OPS1, params.b_g1[nb_in:], b_aux_sc))
    b2_ans = add2(msm(OPS2, params.b_g2[:nb_in], b_in_sc), msm(OPS2, params.b_g2[nb_in:], b_aux_sc))

    if vk.delta_g1 is None or vk.delta_g2 is None:
        raise ValueError("UnexpectedIdentity")
    g_a = add1(bn.jac_mul(OPS1, J1(vk.delta_g1), r), J1(vk.alpha_g1))
    g_b = add2(bn.jac_mul(OPS2, J2(vk.delta_g2), s), J2(vk.beta_g2))
    g_c = bn.jac_mul(OPS1, J1(vk.delta_g1), r * s % R)
    g_c = add1(g_c, bn.jac_mul(OPS1, J1(vk.alpha_g1), s))
    g_c = add1(g_c, bn.jac_mul(OPS1, J1(vk.beta_g1), r))
    g_a = add1(g_a, a_ans)
    g_c = add1(g_c, bn.jac_mul(OPS1, a_ans, s))
    g_b = add2(g_b, b2_ans)
    g_c = add1(g_c, bn.jac_mul(OPS1, b1_ans, r))
    g_c = add1(g_c, h_acc)
    g_c = add1(g_c, l_acc)
    proof = Proof(bn.to_affine(OPS1, g_a), bn.to_affine(OPS2, g_b), bn.to_affine(OPS1, g_c))
    return (proof, h) if return_h else proof


def prove_scalar_side(params: Params, gates, inputs, aux, r, s):
    """Trapdoor shortcut: the discrete logs of (A, B, C) w.r.t. g1/g2.
    Needs params.dlog (oracle-only).  O(n) field work at any size."""
    asg = evaluate_r1cs(gates, inputs, aux)
    h = h_coefficients(asg.a, asg.b, asg.c)
    d = params.dlog
    w = inputs + aux
    dot = lambda xs, ys: sum(x * y for x, y in zip(xs, ys)) % R
    a_sum = dot(d["a_full"], w)
    b_sum = dot(d["b_full"], w)
    A = (d["alpha"] + a_sum + r * d["delta"]) % R
    B = (d["beta"] + b_sum + s * d["delta"]) % R
    C = (dot(d["h"], h) + dot(d["l"], aux) + s * A + r * B - r * s % R * d["delta"]) % R
    return A, B, C


# --------------------------------------------------------------------------
# Verifier (bellman verify_proof restated, App. C.6)
# --------------------------------------------------------------------------
def verify(vk: VerifyingKey, proof: Proof, public_inputs: List[int]) -> bool:
    if len(public_inputs) + 1 != len(vk.ic):
        raise ValueError("MalformedVerifyingKey")
    acc = bn.to_jac(OPS1, vk.ic[0])
    for x, pt in zip(public_inputs, vk.ic[1:]):
        acc = bn.jac_add(OPS1, acc, bn.jac_mul(OPS1, bn.to_jac(OPS1, pt), x))
    ic = bn.to_affine(OPS1, acc)
    neg = lambda p: bn.pt_neg(OPS1, p)
    # e(A,B) * e(-IC,gamma) * e(-C,delta) * e(-alpha,beta) == 1
    return bn.pairing_product_is_one([
        (proof.a, proof.b), (neg(ic), vk.gamma_g2), (neg(proof.c), vk.delta_g2),
        (neg(vk.alpha_g1), vk.beta_g2)])
