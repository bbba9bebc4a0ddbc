"""Regenerate tests/golden/*.json from the Python oracle (run in this container)."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import codec, groth16 as og, synth  # noqa: E402

n_rows, seed = 40, synth.SEED_BASE + 4040
gates, inp, aux = synth.synth_circuit(n_rows, seed)
td, r, s = synth.synth_trapdoor(seed)
P = og.setup(gates, 2, len(aux), td)
proof, h = og.prove(P, gates, inp, aux, r, s, return_h=True)
assert og.verify(P.vk, proof, inp[1:])
pb = codec.bellman_params_bytes(P)
raw = b"".join(codec.gate_borsh(g) for g in gates)
out = {"n_rows": n_rows, "seed": seed, "aux_first4": [hex(x) for x in aux[:4]], "h_first4": [hex(x) for x in h[:4]],
       "proof_raw_hex": codec.proof_raw(proof).hex(), "proof_borsh_hex": codec.proof_borsh(proof).hex(),
       "params_sha256": hashlib.sha256(pb).hexdigest(), "r": hex(r), "s": hex(s),
       "inputs": [hex(x) for x in inp], "aux": [hex(x) for x in aux],
       "gates_raw_hex": raw.hex(), "bellman_params_hex": pb.hex()}
d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(d, exist_ok=True)
json.dump(out, open(os.path.join(d, "synth_rows40.json"), "w"))
print("wrote", len(json.dumps(out)), "bytes")
