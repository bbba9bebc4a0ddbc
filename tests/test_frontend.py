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


# ---------------------------------------------------------------- configs[1]: EdDSA-Poseidon ---
def _rand_subgroup_point(jj, rng):
    """EdwardsPoint::rand (native/ecc.rs:93-100) times the cofactor."""
    while True:
        y = rng.randrange(bn.R)
        y2 = y * y % bn.R
        x = fe.fr_sqrt((y2 - 1) * pow(jj.d * y2 + 1, -1, bn.R) % bn.R)
        if x is not None:
            return fe.ed_mul((x, y), 8, jj.d)


def test_jubjub_params_and_generator():
    jj = fe.JubJubBN256()
    # engines/bn256/mod.rs:50-56: the Montgomery form of the same curve, u a non-residue
    assert jj.a == 168698 and jj.b * (1 + jj.d) % bn.R == bn.R - 4
    assert fe.fr_sqrt(jj.u) is None
    assert jj.in_curve(jj.g) and jj.g != (0, 1)
    assert fe.ed_mul(jj.g, fe.FS, jj.d) == (0, 1)                   # prime-order subgroup
    assert fe.FS_BITS == 251 and fe.R_BITS == 254
    # Montgomery <-> Edwards round trip, decompress recovers the point from x
    assert fe.mont_into_edwards(fe.ed_into_montgomery(jj.g)) == jj.g
    assert jj.subgroup_decompress(jj.g[0]) == jj.g
    p = fe.ed_mul(jj.g, 12345, jj.d)
    assert fe.ed_add(p, jj.g, jj.d) == fe.ed_mul(jj.g, 12346, jj.d)


def test_eddsa_native_sign_verify():
    jj, P = fe.JubJubBN256(), fe.PoseidonParams(4, 8, 54)
    rng = random.Random(21)
    sk, m = rng.randrange(fe.FS), rng.randrange(bn.R)
    s, r = fe.eddsaposeidon_sign(sk, m, P, jj)
    a = fe.ed_mul(jj.g, sk, jj.d)[0]
    assert fe.eddsaposeidon_verify(s, r, a, m, P, jj)
    assert not fe.eddsaposeidon_verify(s, r, a, (m + 1) % bn.R, P, jj)
    assert not fe.eddsaposeidon_verify((s + 1) % fe.FS, r, a, m, P, jj)


def test_ecmul_gadgets_cost_what_the_readme_says():
    """README.md:50-51: ecmul 254 bits = 2,296 constraints, ecmul_const 254 bits = 513 (the shapes of
    tests/circuit_ecc.rs:153-201); both must also compute the native product."""
    jj = fe.JubJubBN256()
    rng = random.Random(22)
    p, n = _rand_subgroup_point(jj, rng), rng.randrange(bn.R)
    want = fe.ed_mul(p, n % fe.FS, jj.d)
    for const_base, cost in ((False, 2296), (True, 513)):
        cs = fe.BuildCS()
        sp = fe.CEdwardsPoint.from_const(cs, p) if const_base else fe.CEdwardsPoint.alloc(cs, p)
        bits = fe.c_into_bits_le_strict(cs.alloc(n))
        assert [b.value for b in bits] == [(n >> i) & 1 for i in range(254)]
        g0 = len(cs.gates)
        res = sp.mul(bits, jj)
        assert len(cs.gates) - g0 == cost
        assert (res.x.value, res.y.value) == want


def test_bit_gadgets():
    rng = random.Random(23)
    for limit, v, ct in ((8, 200, 199), (8, 200, 200), (9, 200, 201), (251, fe.FS - 1, fe.FS - 1), (251, fe.FS, fe.FS - 1)):
        cs = fe.BuildCS()
        bits = fe.c_into_bits_le(cs.alloc(v), limit)
        assert [b.value for b in bits] == [(v >> i) & 1 for i in range(limit)]
        assert fe.c_comp_constant(bits, ct).value == (1 if v > ct else 0)
    cs = fe.BuildCS()
    x = cs.alloc(rng.randrange(1, bn.R))
    assert fe.c_is_zero(x).value == 0 and fe.c_is_zero(x - x).value == 1 and fe.c_is_zero(cs.alloc(0)).value == 1
    q = fe.c_div_unchecked(x, cs.alloc(7))
    assert q.value * 7 % bn.R == x.value
    cols = [[rng.randrange(bn.R) for _ in range(8)] for _ in range(2)]
    for idx in range(8):
        cs = fe.BuildCS()
        s = [fe.c_alloc_bool(cs, (idx >> j) & 1) for j in range(3)]
        assert [o.value for o in fe.c_mux3(s, cols)] == [cols[0][idx], cols[1][idx]]


def test_eddsa_circuit_shape():
    """configs[1].  The gadget code at the reference's commit gives 4,121 gates for c_eddsaposeidon_verify
    (2 x 20 subgroup_decompress + 255 Poseidon + 510 strict bits + 2,296 ecmul + 251 + 253 range-checked
    s bits + 507 fixed-base mul + 6 add + 3 is_zero); README.md:53 quotes 3,860 (an older gadget set: its
    19-constraint 'oncurve+subgroup check' row is 25 here too).  With inputize + assert_const + the two
    bellman input rows: 4,125 rows -> m = 2^13."""
    rng = random.Random(24)
    sk, m = rng.randrange(fe.FS), rng.randrange(bn.R)
    gates, inputs, aux = fe.eddsa_circuit(sk, m)
    assert len(gates) == 4121 + 2 and inputs == [1, m]
    w = {(INPUT, i): v for i, v in enumerate(inputs)}
    w.update({(AUX, i): v for i, v in enumerate(aux)})
    for A, B, C in gates:
        ev = [sum(c * w[k] for c, k in lc) % bn.R for lc in (A, B, C)]
        assert ev[0] * ev[1] % bn.R == ev[2]
        for lc in (A, B, C):
            assert [k for _, k in lc] == sorted(k for _, k in lc) and all(c % bn.R for c, _ in lc)
    # the same circuit for another key and message: identical gate structure (only the witness moves)
    gates2, _, aux2 = fe.eddsa_circuit(rng.randrange(fe.FS), rng.randrange(bn.R))
    assert gates2 == gates and aux2 != aux


def test_frontend_oracle_chain_matches_committed_golden():
    """tests/golden/frontend_circuits.json (tools/gen_golden_frontend.py): front end -> Python setup (fixed
    trapdoor) -> C++ prove (fixed r, s) -> independent pairing check, for the EdDSA circuit; the digests of the
    gate stream and of the bellman Parameters and the 256 proof bytes must be the committed ones."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import gen_golden_frontend as gg
    golden = json.load(open(os.path.join(root, "tests", "golden", "frontend_circuits.json")))
    assert set(golden) == {"cfg1_poseidon_merkle", "cfg2_eddsa_poseidon"}
    got = gg.oracle_chain("cfg2_eddsa_poseidon")
    assert got == golden["cfg2_eddsa_poseidon"]
    assert got["verifies"] and got["wrong_input_rejected"]
    # cfg 1: the cheap half (front end only; its full chain runs in the generator)
    import hashlib
    from oracle import codec
    gates, inp, aux, *_ = gg.build_case("cfg1_poseidon_merkle")
    g1 = golden["cfg1_poseidon_merkle"]
    assert hashlib.sha256(b"".join(codec.gate_borsh(g) for g in gates)).hexdigest() == g1["gates_sha256"]
    assert hex(inp[1]) == g1["public_input"] and len(aux) == g1["n_aux"]
