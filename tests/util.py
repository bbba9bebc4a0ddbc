"""Shared helpers for the parity tests (oracle <-> C ABI)."""
import random

import numpy as np

from oracle import bn254 as bn
from oracle import codec


def fr_np(vals):
    """canonical ints -> uint64[n,4] Montgomery (Num<Fr> in memory)."""
    buf = b"".join(codec.fr_raw(v) for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def fq_np(vals):
    buf = b"".join(codec.fq_raw(v) for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def fr_list(arr):
    b = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [codec.fr_unraw(b[i:i + 32]) for i in range(0, len(b), 32)]


def fq_list(arr):
    b = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [codec.fq_unraw(b[i:i + 32]) for i in range(0, len(b), 32)]


def edge_values(mod, rng, n):
    vals = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, pow(2, 255, mod), (mod - 1) // 2, 0xFFFFFFFF, 1 << 32,
            (1 << 64) - 1, 1 << 64, (1 << 128) - 1, (1 << 253) + 1]
    while len(vals) < n:
        vals.append(rng.randrange(mod))
    return vals[:n]


def random_g1(rng, n):
    fb = getattr(random_g1, "_fb", None)
    if fb is None:
        fb = random_g1._fb = bn.FixedBase(bn.OPS1, bn.G1_GEN)
    ks = [rng.randrange(1, bn.R) for _ in range(n)]
    return ks, fb.mul_many(ks)


def random_g2(rng, n):
    fb = getattr(random_g2, "_fb", None)
    if fb is None:
        fb = random_g2._fb = bn.FixedBase(bn.OPS2, bn.G2_GEN)
    ks = [rng.randrange(1, bn.R) for _ in range(n)]
    return ks, fb.mul_many(ks)


def random_gate_blob(n_gates, n_in, n_aux, seed, terms=(3, 3, 1), pool=300):
    """A borsh gate stream (cs.rs:184-223 framing: per LC a u32 count, then 37-byte terms) with random content,
    built with numpy: coefficients are 1 (50 %), r-1 (10 %), one of `pool` fixed values (20 %) or a fresh
    253-bit number (20 %); variables uniform over inputs and aux.  Not a satisfiable circuit -- parser food."""
    rng = np.random.default_rng(seed)
    per_gate = sum(4 + 37 * t for t in terms)
    raw = np.zeros((n_gates, per_gate), dtype=np.uint8)
    poolv = rng.integers(0, 256, size=(pool, 32), dtype=np.uint8)
    poolv[:, 31] &= 0x1F
    one = np.zeros(32, dtype=np.uint8); one[0] = 1
    mone = np.frombuffer((bn.R - 1).to_bytes(32, "little"), dtype=np.uint8)
    pos = 0
    for t in terms:
        raw[:, pos:pos + 4] = np.frombuffer(np.uint32(t).tobytes(), dtype=np.uint8)
        pos += 4
        for _ in range(t):
            kind = rng.random(n_gates)
            coef = rng.integers(0, 256, size=(n_gates, 32), dtype=np.uint8)
            coef[:, 31] &= 0x1F
            sel = kind < 0.5
            coef[sel] = one
            sel = (kind >= 0.5) & (kind < 0.6)
            coef[sel] = mone
            sel = (kind >= 0.6) & (kind < 0.8)
            coef[sel] = poolv[rng.integers(0, pool, size=int(sel.sum()))]
            raw[:, pos:pos + 32] = coef
            tag = (rng.random(n_gates) < 0.9).astype(np.uint8)
            raw[:, pos + 32] = tag
            idx = np.where(tag == 1, rng.integers(0, n_aux, size=n_gates), rng.integers(0, n_in, size=n_gates)).astype("<u4")
            raw[:, pos + 33:pos + 37] = idx.view(np.uint8).reshape(-1, 4)
            pos += 37
    return raw.tobytes()
