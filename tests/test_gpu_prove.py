"""End-to-end parity: setup / prove / verify through the C ABI vs the oracle (byte-exact)."""
import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import codec
from oracle import groth16 as og
from oracle import synth
from tests.util import fr_np, fr_list

pytestmark = pytest.mark.gpu


def oracle_case(n_rows, seed):
    gates, inp, aux = synth.synth_circuit(n_rows, seed)
    td, r, s = synth.synth_trapdoor(seed)
    P = og.setup(gates, 2, len(aux), td)
    return gates, inp, aux, td, r, s, P


@pytest.mark.parametrize("n_rows", [3, 17, 64, 700])
def test_prove_matches_oracle_bytes(ctx, n_rows):
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 100 + n_rows
    gates, inp, aux, td, r, s, P = oracle_case(n_rows, seed)
    ref_proof, ref_h = og.prove(P, gates, inp, aux, r, s, return_h=True)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
    # through the reference's own byte framing (Parameters::write -> read)
    params = fb.Parameters.read(params.write())
    inputs, proof, h = fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux), r, s, ctx, return_h=True)
    assert fr_list(h) == ref_h
    assert proof.to_raw() == codec.proof_raw(ref_proof)
    assert proof.serialize() == codec.proof_borsh(ref_proof)
    assert fr_list(inputs) == inp[1:]
    assert fb.verify(params.get_vk(), proof, inputs)
    assert og.verify(P.vk, codec.proof_unraw(proof.to_raw()), inp[1:])
    # tampered public input must fail
    bad = fr_np([(inp[1] + 1) % bn.R])
    assert not fb.verify(params.get_vk(), proof, bad)
    params.unload()


@pytest.mark.parametrize("n_rows", [5, 64, 300])
def test_setup_matches_oracle_bytes(ctx, n_rows):
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 200 + n_rows
    gates, inp, aux, td, r, s, P = oracle_case(n_rows, seed)
    circ = fb.Circuit.synthetic(n_rows, seed)
    params = fb.setup(circ, ctx, trapdoor=[td.alpha, td.beta, td.gamma, td.delta, td.tau])
    assert params.bellman_bytes == codec.bellman_params_bytes(P)


def test_prove_synthetic_2_16_scalar_identity_and_verify(ctx):
    """2^16 rows: too big for the Python prover, so check (i) the trapdoor scalar identity
    A == (alpha + A(tau) + r delta) G1 etc. with the oracle's O(n) scalar side, (ii) verify."""
    import fawkes_crypto_b200 as fb
    n_rows = 1 << 16
    seed = synth.SEED_BASE + 16
    circ = fb.Circuit.synthetic(n_rows, seed)
    tdn = np.zeros((7, 4), dtype=np.uint64)
    fb.native.check(fb.native.lib.fb_synth_trapdoor(seed, tdn.ctypes.data))
    td = fr_list(tdn)
    params = fb.setup(circ, ctx, trapdoor=td[:5])
    wi, wa = circ.witness()
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, td[5], td[6], ctx)
    assert fb.verify(params.get_vk(), proof, inputs)
    gates, inp, aux = synth.synth_circuit(n_rows, seed)
    assert fr_list(wa) == aux
    P = og.setup(gates, 2, len(aux), og.Trapdoor(*td[:5]), want_points=False)
    A, B, C = og.prove_scalar_side(P, gates, inp, aux, td[5], td[6])
    pr = codec.proof_unraw(proof.to_raw())
    assert pr.a == bn.pt_mul(bn.OPS1, bn.G1_GEN, A)
    assert pr.b == bn.pt_mul(bn.OPS2, bn.G2_GEN, B)
    assert pr.c == bn.pt_mul(bn.OPS1, bn.G1_GEN, C)
    params.unload()


def test_error_paths(ctx):
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 300
    gates, inp, aux, td, r, s, P = oracle_case(20, seed)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    good = codec.bellman_params_bytes(P)
    # wrong witness length
    params = fb.Parameters(good, len(gates), codec.brotli_compress(raw))
    with pytest.raises(fb.native.FbError) as e:
        fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux[:-1]), r, s, ctx)
    assert e.value.code == -1
    params.unload()
    # truncated parameters
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters(good[:-10], len(gates), codec.brotli_compress(raw)).load(ctx)
    assert e.value.code == -3
    # point not on curve (checked = True)
    bad = bytearray(good)
    bad[64 + 63] ^= 1   # beta_g1.y
    with pytest.raises(fb.native.FbError):
        fb.Parameters(bytes(bad), len(gates), codec.brotli_compress(raw)).load(ctx)
    # gate count mismatch
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters(good, len(gates) + 1, codec.brotli_compress(raw)).load(ctx)
    assert e.value.code == -3
