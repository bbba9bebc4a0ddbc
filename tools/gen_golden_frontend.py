"""Regenerate tests/golden/frontend_circuits.json (run in this container, CPU only): for configs[0] (Poseidon
Merkle proof, depth 32) and configs[1] (EdDSA-Poseidon signature) the front-end restatement builds the R1CS and
the witness from fixed seeds, the Python oracle runs setup with the fixed trapdoor, the C++ restatement proves
with fixed r, s, and the Python oracle's independent pairing check accepts the proof.  The fixture pins the
digests of the gate stream, of the bellman Parameters bytes and the 256 proof bytes."""
import hashlib
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bn254 as bn, codec, cpu, groth16 as og, synth  # noqa: E402
from oracle import frontend as fe  # noqa: E402
from tests.util import fr_np  # noqa: E402


def build_case(name):
    if name == "cfg1_poseidon_merkle":
        rng = random.Random(2026)
        leaf = rng.randrange(bn.R)
        sibling = [rng.randrange(bn.R) for _ in range(32)]
        path = [rng.random() < 0.5 for _ in range(32)]
        gates, inp, aux = fe.merkle_circuit(leaf, sibling, path)
        seed = synth.SEED_BASE + 1
    else:
        rng = random.Random(2027)
        gates, inp, aux = fe.eddsa_circuit(rng.randrange(fe.FS), rng.randrange(bn.R))
        seed = synth.SEED_BASE + 2
    td, r, s = synth.synth_trapdoor(seed)
    return gates, inp, aux, td, r, s


def csr_of(gates, n_in):
    """(rowptr, col, coef) per matrix for oracle/cpu.py: uint32, uint32, uint64[nnz,4] Montgomery."""
    rps, cols, coefs = [], [], []
    for m in range(3):
        rp, cl, cf = [0], [], []
        for g in gates:
            for c, (tag, idx) in g[m]:
                cl.append(idx if tag == 0 else n_in + idx)
                cf.append(c)
            rp.append(len(cl))
        rps.append(np.array(rp, dtype=np.uint32))
        cols.append(np.array(cl, dtype=np.uint32))
        coefs.append(fr_np(cf) if cf else np.zeros((0, 4), dtype=np.uint64))
    return rps, cols, coefs


def oracle_chain(name):
    gates, inp, aux, td, r, s = build_case(name)
    P = og.setup(gates, len(inp), len(aux), td)
    pb = codec.bellman_params_bytes(P)
    rp, cl, cf = csr_of(gates, len(inp))
    proof, _, _ = cpu.prove(pb, len(gates), len(inp), len(aux), rp, cl, cf, fr_np(inp), fr_np(aux),
                            np.frombuffer(codec.fr_raw(r), dtype=np.uint64), np.frombuffer(codec.fr_raw(s), dtype=np.uint64),
                            min(cpu.hw_threads(), 8))
    ok = og.verify(P.vk, codec.proof_unraw(proof), inp[1:])
    bad = og.verify(P.vk, codec.proof_unraw(proof), [(inp[1] + 1) % bn.R])
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    return {"n_gates": len(gates), "n_in": len(inp), "n_aux": len(aux), "nnz": sum(len(x) for g in gates for x in g),
            "gates_sha256": hashlib.sha256(raw).hexdigest(), "params_sha256": hashlib.sha256(pb).hexdigest(),
            "public_input": hex(inp[1]), "proof_raw_hex": bytes(proof).hex(), "verifies": bool(ok),
            "wrong_input_rejected": not bad}


if __name__ == "__main__":
    out = {n: oracle_chain(n) for n in ("cfg1_poseidon_merkle", "cfg2_eddsa_poseidon")}
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    json.dump(out, open(os.path.join(d, "frontend_circuits.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
