"""configs[0] on the GPU path: setup + prove + verify of the depth-32 Poseidon Merkle-proof circuit (7,362 gates,
m = 2^13) built by oracle/frontend.py, fixed trapdoor and r, s.  Prints one JSON line."""
import json, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
from oracle import bn254 as bn, codec, synth
from oracle import frontend as fe
from tests.util import fr_np

ctx = fb.Context(0)
rng = random.Random(2026)
leaf = rng.randrange(bn.R); sibling = [rng.randrange(bn.R) for _ in range(32)]; path = [rng.random() < 0.5 for _ in range(32)]
t = time.perf_counter(); gates, inp, aux = fe.merkle_circuit(leaf, sibling, path); t_front = time.perf_counter() - t
raw = b"".join(codec.gate_borsh(g) for g in gates)
circ = fb.Circuit.from_raw_gates(raw, len(gates), 2, len(aux))
td, r, s = synth.synth_trapdoor(synth.SEED_BASE + 1)
t = time.perf_counter()
params = fb.setup(circ, ctx, trapdoor=[td.alpha, td.beta, td.gamma, td.delta, td.tau], gates_blob=codec.brotli_compress(raw))
t_setup = time.perf_counter() - t
wi, wa = fr_np(inp), fr_np(aux)
t = time.perf_counter(); params.load(ctx); t_load = time.perf_counter() - t
for _ in range(5):
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, r, s, ctx)
n = 50
t = time.perf_counter()
for _ in range(n):
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, r, s, ctx)
t_prove = (time.perf_counter() - t) / n
t = time.perf_counter(); ok = fb.verify(params.get_vk(), proof, inputs); t_verify = time.perf_counter() - t
print(json.dumps({"config": "poseidon merkle proof depth 32 (7,328 + 34 gates, 7,364 rows, m = 2^13), fixed trapdoor and r,s",
                  "nnz": circ.shape()["nnz"], "setup_s": t_setup, "key_load_s": t_load, "prove_ms": t_prove * 1e3,
                  "prove_stage_ms": params.timings(), "verify_ms": t_verify * 1e3, "verifies": bool(ok),
                  "front_end_python_s": t_front, "msm": {k: params.info()[k] for k in ("msm_window_bits", "msm_windows", "msm_tables")}}))
