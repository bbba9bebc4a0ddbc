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
