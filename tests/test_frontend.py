"""configs[0] front end (oracle/frontend.py): primitives against published vectors, circuit shape against
the reference's README (7,328 constraints for the depth-32 Poseidon Merkle proof)."""
import random

from oracle import bn254 as bn
from oracle import frontend as fe
from oracle.groth16 import INPUT, AUX


def test_keccak256_known_answers():
    assert fe.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert fe.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    # more than one block (rate 136)
    assert len(fe.keccak256(b"x" * 300)) == 32 and fe.keccak256(b"x" * 300) != fe.keccak256(b"x" * 301)


def test_chacha20_block_rfc7539_vector():
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    out = fe.chacha20_block(key, 1, 0x09000000, 0x4A000000, 0)
    assert out == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                   0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def test_poseidon_params_shape_and_mds():
    P = fe.PoseidonParams(3, 8, 53)
    assert len(P.c) == 61 and all(len(r) == 3 for r in P.c)
    assert all(0 <= v < bn.R for r in P.c for v in r)
    # m is a Cauchy matrix: every 2x2 minor is non-zero (invertible mixing layer)
    for i in range(3):
        for j in range(i + 1, 3):
            for k in range(3):
                for l in range(k + 1, 3):
                    assert (P.m[i][k] * P.m[j][l] - P.m[i][l] * P.m[j][k]) % bn.R != 0
    assert fe.PoseidonParams(3, 8, 53).c == P.c                      # deterministic
    assert fe.PoseidonParams(3, 8, 53, salt="x").c != P.c


def test_poseidon_circuit_matches_native_and_costs_228_gates():
    P = fe.PoseidonParams(3, 8, 53)
    rng = random.Random(5)
    a, b = rng.randrange(bn.R), rng.randrange(bn.R)
    cs = fe.BuildCS()
    out = fe.c_poseidon([cs.alloc(a), cs.alloc(b)], P)
    assert out.value == fe.poseidon([a, b], P)
    assert len(cs.gates) == 8 * 3 * 3 + 53 * 3 - 3 == 228          # SURVEY App. E / README.md:52


def test_poseidon_4_8_54_costs_255_gates():
    """The other Poseidon row of the reference's benchmark table (README.md:48): poseidon hash (4, 8, 54) = 255."""
    P = fe.PoseidonParams(4, 8, 54)
    rng = random.Random(6)
    vals = [rng.randrange(bn.R) for _ in range(3)]
    cs = fe.BuildCS()
    out = fe.c_poseidon([cs.alloc(v) for v in vals], P)
    assert out.value == fe.poseidon(vals, P)
    assert len(cs.gates) == 255


def test_merkle_circuit_shape_is_cfg1():
    rng = random.Random(11)
    leaf = rng.randrange(bn.R)
    sibling = [rng.randrange(bn.R) for _ in range(32)]
    path = [rng.random() < 0.5 for _ in range(32)]
    gates, inputs, aux = fe.merkle_circuit(leaf, sibling, path)
    assert len(gates) == 7328 + 32 + 1 + 1 == 7362                   # + 2 bellman input rows = 7,364 -> m = 2^13
    assert len(inputs) == 2 and len(aux) == 1 + 1 + 32 + 32 + 7328 == 7394
    assert inputs[1] == fe.poseidon_merkle_proof_root(leaf, sibling, path, fe.PoseidonParams(3, 8, 53))
    # every gate holds on the witness, LCs are sorted Input < Aux with no zero coefficients
    w = {(INPUT, i): v for i, v in enumerate(inputs)}
    w.update({(AUX, i): v for i, v in enumerate(aux)})
    widest = 0
    for A, B, C in gates:
        ev = [sum(c * w[k] for c, k in lc) % bn.R for lc in (A, B, C)]
        assert ev[0] * ev[1] % bn.R == ev[2]
        for lc in (A, B, C):
            assert [k for _, k in lc] == sorted(k for _, k in lc) and all(c % bn.R for c, _ in lc)
            widest = max(widest, len(lc))
    assert widest > 50          # lanes 1-2 accumulate through the 53 partial rounds
    # first gates: inputize, then the 32 path-bit checks
    assert gates[0] == ([(1, (AUX, 0))], [(1, (INPUT, 0))], [(1, (INPUT, 1))])
    assert gates[1][0] == [(1, (AUX, 34))] and gates[1][2] == []
