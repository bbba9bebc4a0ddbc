"""Regenerate tests/golden/bench_cfgs.npz (CPU only, a few minutes): the inputs and expected outputs bench.py needs to
run BASELINE.json configs[0] and configs[1] on the GPU without importing the oracle's front end.

  cfg1  Poseidon Merkle proof, depth 32 (tests/bellman_groth16.rs:19-47): gate stream (brotli), witness, trapdoor,
        r, s, expected Parameters digest and proof bytes (== tests/golden/frontend_circuits.json).
  cfg2  EdDSA-Poseidon signature circuit (circuit/eddsaposeidon.rs:16-47), 256 proofs per run: ONE gate stream,
        8 distinct signatures (keys, messages, witnesses) cycled over the 256 slots, r_i = r0 + 5 i, s_i = s0 + 9 i;
        expected: the 256 proofs of the C++ CPU oracle under the C++ oracle's own setup.
Circuits and witnesses come from oracle/frontend.py; setup and proofs from oracle/cpu_setup.cpp / cpu_prover.cpp."""
import hashlib
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import bn254 as bn, codec, cpu, synth  # noqa: E402
from oracle import frontend as fe  # noqa: E402
from tests.util import fr_np  # noqa: E402
import gen_golden_frontend as gg  # noqa: E402

COUNT, DISTINCT = 256, 8


def main():
    out = {}
    front = json.load(open(os.path.join(ROOT, "tests", "golden", "frontend_circuits.json")))
    th = min(cpu.hw_threads(), 8)
    # ---- cfg1
    gates, inp, aux, td, r, s = gg.build_case("cfg1_poseidon_merkle")
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    rp, cl, cf = gg.csr_of(gates, len(inp))
    circ = cpu.Circuit.from_csr(len(gates), len(inp), len(aux), rp, cl, cf)
    tdm = fr_np([td.alpha, td.beta, td.gamma, td.delta, td.tau, r, s])
    pbuf, _ = cpu.setup(circ, tdm, th)
    assert hashlib.sha256(pbuf.array).hexdigest() == front["cfg1_poseidon_merkle"]["params_sha256"]
    proof, _, _ = cpu.prove_circuit(pbuf, circ, tdm[5], tdm[6], th, inputs=fr_np(inp), aux=fr_np(aux))
    assert proof.hex() == front["cfg1_poseidon_merkle"]["proof_raw_hex"]
    out["cfg1_gates_brotli"] = np.frombuffer(codec.brotli_compress(raw), dtype=np.uint8)
    out["cfg1_shape"] = np.array([len(gates), len(inp), len(aux)], dtype=np.uint32)
    out["cfg1_inputs"], out["cfg1_aux"] = fr_np(inp), fr_np(aux)
    out["cfg1_trapdoor_r_s"] = tdm
    out["cfg1_params_sha256"] = np.frombuffer(hashlib.sha256(pbuf.array).digest(), dtype=np.uint8)
    out["cfg1_proof"] = np.frombuffer(proof, dtype=np.uint8)
    print("cfg1", len(gates), "gates", len(raw), "raw bytes", out["cfg1_gates_brotli"].size, "brotli", flush=True)
    # ---- cfg2
    rng = random.Random(2027)
    jj, P = fe.JubJubBN256(), fe.PoseidonParams(4, 8, 54)
    cases = [fe.eddsa_circuit(rng.randrange(fe.FS), rng.randrange(bn.R), P, jj) for _ in range(DISTINCT)]
    gates = cases[0][0]
    assert all(c[0] == gates for c in cases), "the gate list must not depend on key or message"
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    assert hashlib.sha256(raw).hexdigest() == front["cfg2_eddsa_poseidon"]["gates_sha256"]
    n_in, n_aux = len(cases[0][1]), len(cases[0][2])
    rp, cl, cf = gg.csr_of(gates, n_in)
    circ = cpu.Circuit.from_csr(len(gates), n_in, n_aux, rp, cl, cf)
    td, r0, s0 = synth.synth_trapdoor(synth.SEED_BASE + 2)
    tdm = fr_np([td.alpha, td.beta, td.gamma, td.delta, td.tau, r0, s0])
    pbuf, _ = cpu.setup(circ, tdm, th)
    assert hashlib.sha256(pbuf.array).hexdigest() == front["cfg2_eddsa_poseidon"]["params_sha256"]
    wit_in = np.stack([fr_np(c[1]) for c in cases])
    wit_aux = np.stack([fr_np(c[2]) for c in cases])
    rs = [(r0 + 5 * i) % bn.R for i in range(COUNT)]
    ss = [(s0 + 9 * i) % bn.R for i in range(COUNT)]
    rsm, ssm = fr_np(rs), fr_np(ss)
    proofs = np.zeros((COUNT, 256), dtype=np.uint8)
    for i in range(COUNT):
        pr, _, _ = cpu.prove_circuit(pbuf, circ, rsm[i], ssm[i], 1, inputs=wit_in[i % DISTINCT], aux=wit_aux[i % DISTINCT])
        proofs[i] = np.frombuffer(pr, dtype=np.uint8)
        if i % 32 == 0:
            print("cfg2 proof", i, flush=True)
    assert proofs[0].tobytes().hex() == front["cfg2_eddsa_poseidon"]["proof_raw_hex"]
    out["cfg2_gates_brotli"] = np.frombuffer(codec.brotli_compress(raw), dtype=np.uint8)
    out["cfg2_shape"] = np.array([len(gates), n_in, n_aux], dtype=np.uint32)
    out["cfg2_inputs"], out["cfg2_aux"] = wit_in, wit_aux
    out["cfg2_trapdoor_r_s"] = tdm
    out["cfg2_params_sha256"] = np.frombuffer(hashlib.sha256(pbuf.array).digest(), dtype=np.uint8)
    out["cfg2_proofs_sha256"] = np.frombuffer(hashlib.sha256(proofs.tobytes()).digest(), dtype=np.uint8)
    out["cfg2_proofs_first8"] = proofs[:8].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bench_cfgs.npz"), **out)
    print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", "bench_cfgs.npz")), "bytes")


if __name__ == "__main__":
    main()
