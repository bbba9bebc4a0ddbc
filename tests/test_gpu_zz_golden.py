"""Last GPU test file in collection order: the committed golden fixture of the two real circuits
(tests/golden/frontend_circuits.json, made on the CPU by tools/gen_golden_frontend.py: Python setup + C++ prove)
against GPU setup + GPU prove.  Parameters digest and proof bytes must match."""
import hashlib
import json
import os
import sys

import pytest

from oracle import codec
from tests.util import fr_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["cfg2_eddsa_poseidon", "cfg1_poseidon_merkle"])
def test_gpu_setup_and_prove_equal_committed_golden(ctx, name):
    import fawkes_crypto_b200 as fb
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_golden_frontend as gg
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "frontend_circuits.json")))[name]
    gates, inp, aux, td, r, s = gg.build_case(name)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    assert hashlib.sha256(raw).hexdigest() == golden["gates_sha256"]
    circ = fb.Circuit.from_raw_gates(raw, len(gates), len(inp), len(aux), ctx=ctx)
    params = fb.setup(circ, ctx, trapdoor=[td.alpha, td.beta, td.gamma, td.delta, td.tau])
    assert hashlib.sha256(bytes(params.bellman_bytes)).hexdigest() == golden["params_sha256"]
    inputs, proof = fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux), r, s, ctx)
    assert proof.to_raw().hex() == golden["proof_raw_hex"]
    assert fb.verify(params.get_vk(), proof, inputs)
    params.unload()
