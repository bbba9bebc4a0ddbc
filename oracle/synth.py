"""Synthetic random R1CS of SURVEY.md section 8(d) -- Python restatement.

TEST INFRASTRUCTURE ONLY.  The same generator exists in C++
(fawkes-crypto_b200/csrc/synth.cpp, exported as fb_synth_*) for the benchmark
sizes; tests check the two agree bit for bit.

Shape (mimics CNum::mul_assign, fawkes-crypto/src/circuit/r1cs/num.rs:253-272,
and BuildCS::inputize, circuit/r1cs/cs.rs:309-318):
  n_in = 2 (ONE + one public), n_gates = n_rows - 2.
  16 aux values sampled directly; inputs[1] = aux[0].
  row 0: [1*Aux0] * [1*Input0] = [1*Input1]                    (inputize)
  row i: A_i, B_i = 3 terms each over uniformly random earlier variables,
         coeff = 1 w.p. 1/2 else uniform Fr;  new aux = <A,w>*<B,w>;  C_i = 1*new.
PRNG SplitMix64; Fr sample = 4 words LE, top limb masked to 62 bits, reject >= r.
"""
from __future__ import annotations

from .bn254 import R, MASK64
from .groth16 import INPUT, AUX, Trapdoor

SEED_BASE = 0xFA3CE50000
N_INIT_AUX = 16


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def fr(self):
        while True:
            l = [self.next() for _ in range(4)]
            l[3] &= (1 << 62) - 1
            v = l[0] | (l[1] << 64) | (l[2] << 128) | (l[3] << 192)
            if v < R:
                return v


def synth_circuit(n_rows: int, seed: int):
    """Returns (gates, inputs, aux) with canonical integer values."""
    assert n_rows >= 3
    rng = SplitMix64(seed)
    n_gates = n_rows - 2
    aux = [rng.fr() for _ in range(N_INIT_AUX)]
    inputs = [1, aux[0]]
    gates = [([(1, (AUX, 0))], [(1, (INPUT, 0))], [(1, (INPUT, 1))])]

    def val(t):
        return inputs[t[1]] if t[0] == INPUT else aux[t[1]]

    for _ in range(1, n_gates):
        lcs = []
        for _side in range(2):
            lc = []
            for _k in range(3):
                u = rng.next() % (2 + len(aux))
                idx = (INPUT, u) if u < 2 else (AUX, u - 2)
                coeff = 1 if rng.next() & 1 else rng.fr()
                lc.append((coeff, idx))
            lcs.append(lc)
        ea = sum(c * val(t) for c, t in lcs[0]) % R
        eb = sum(c * val(t) for c, t in lcs[1]) % R
        aux.append(ea * eb % R)
        gates.append((lcs[0], lcs[1], [(1, (AUX, len(aux) - 1))]))
    return gates, inputs, aux


def synth_trapdoor(seed: int):
    """(Trapdoor, r, s) from the stream seed ^ 0xB11D."""
    rng = SplitMix64(seed ^ 0xB11D)
    vals = [rng.fr() for _ in range(7)]
    return Trapdoor(*vals[:5]), vals[5], vals[6]
